"""Parallel engine: fabric (process group + symmetric heap), dense path,
sparse path and mode routing."""
