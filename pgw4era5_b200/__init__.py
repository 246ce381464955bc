"""B200-native implementation of PGW4ERA5's per-timestep ERA5 modification path."""
__version__ = "0.1.0"
