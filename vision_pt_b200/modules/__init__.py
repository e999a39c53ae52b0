"""Host-side mirror of /root/reference/src/modules for the hot path (same names, argument meaning, errors)."""
