"""Stub: xformers is imported by UCF_VIT/fsdp/building_blocks.py but only used for FusedAttn.FLASH/CK."""
