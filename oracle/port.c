/* placeholder: replaced below */
