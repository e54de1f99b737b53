"""TEST INFRASTRUCTURE -- not part of the product.  See oracle/README.md."""
