"""pffdtd_b200 -- Blackwell-native engine for the simulation step of bsxfun/pffdtd.

    sim_fdtd   the host run loop / CLI (drop-in for `python -m fdtd.sim_fdtd` and fdtd_main_gpu_*.x)
    engine     ctypes binding of libpffdtd_b200.so (include/pffdtd_b200.h); csrc/ holds the sm_100a kernels
    sim_data   the reference's `struct SimData` and its host prep, restated in numpy
    h5lite     the HDF5 subset of the four input files and sim_outs.h5
    shoebox, folder_prep   synthetic inputs and the "gpu folder" transforms
"""
__version__ = "0.1"
