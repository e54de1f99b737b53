"""gala_b200 -- B200-native (sm_100a) engine for gala's orbit-integration hot path.

Mirrors the reference API for that path only: potentials with the ``CPotential`` parameter
layout, ``StaticFrame`` / ``ConstantRotatingFrame``, ``Hamiltonian.integrate_orbit`` with
Leapfrog / Ruth4 / DOPRI853, and ``MockStreamGenerator`` with ``FardalStreamDF``.
All arithmetic runs in hand-written CUDA kernels behind the C ABI of ``include/gala_b200.h``.
"""
from . import _abi
from ._abi import set_devices, get_devices
from .units import galactic, dimensionless, solarsystem, G_GALACTIC, KMS_TO_KPC_MYR
from .potential import *          # noqa: F401,F403
from .frame import (StaticFrame, ConstantRotatingFrame, static_to_constantrotating, constantrotating_to_static,
                    static_to_static)
from .dynamics import PhaseSpacePosition, Orbit, MockStream, peak_to_peak_period, estimate_dt_n_steps, combine
from .integrate import (pinned_empty, parse_time_specification, LeapfrogIntegrator, Ruth4Integrator, DOPRI853Integrator,
                        leapfrog_integrate_hamiltonian, ruth4_integrate_hamiltonian,
                        dop853_integrate_hamiltonian, integrate_extrema, orbit_extrema, orbit_extrema_list)
from .hamiltonian import Hamiltonian
from .mockstream import (BaseStreamDF, FardalStreamDF, StreaklineStreamDF, LagrangeCloudStreamDF, ChenStreamDF,
                         MockStreamGenerator, DirectNBody, mockstream_dop853, mockstream_leapfrog)

from .nonlinear import fast_lyapunov_max, surface_of_section
from . import potential_io as io        # gala's name for it (gala.potential.potential.io); the file is not called io.py so that it can never shadow the standard library's
from .potential_io import load, save

__version__ = "0.2.0"
