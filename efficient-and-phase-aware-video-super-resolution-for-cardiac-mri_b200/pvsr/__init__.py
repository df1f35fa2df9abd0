"""pvsr: Python host side of the B200-native RefineNet path (ctypes over libpvsr.so, include/pvsr.h)."""
