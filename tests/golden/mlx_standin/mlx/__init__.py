"""NumPy stand-in for the `mlx` package -- TEST INFRASTRUCTURE ONLY (see core.py)."""
