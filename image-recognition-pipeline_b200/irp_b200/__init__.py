"""irp_b200 -- host-side plumbing of the B200 outlier-detection stage (ctypes binding, torch custom ops, engine)."""
