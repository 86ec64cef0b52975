"""Stub: misc.py imports torchdata.datapipes (0.9 API, removed in 0.11) for file listing only."""
