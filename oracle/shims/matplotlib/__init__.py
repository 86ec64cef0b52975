"""Stub: quadtree.py / octree.py import matplotlib.pyplot only for drawing helpers."""
