"""Spatially sharded map for large scenes: placeholder filled in below."""
