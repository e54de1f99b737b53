// scf.cuh -- Hernquist-Ostriker SCF basis-function expansion on the device (placeholder until the
// recurrence implementation lands; the host rejects GB_POT_SCF with -11 while GB_HAVE_SCF is 0).
#pragma once
struct PotSCF {
    GB_DEV static void gradient(const double*, const double*, double, double, double, double&, double&, double&) {}
    GB_DEV static double value(const double*, const double*, double, double, double) { return 0.; }
    GB_DEV static double density(const double*, const double*, double, double, double) { return 0.; }
};
