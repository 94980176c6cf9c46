"""Empty stand-in: only AbundanceVectorLocal (out of scope) uses h5py."""
