"""Stub: UCF_VIT/utils/misc.py imports nibabel at module top (file IO only)."""
