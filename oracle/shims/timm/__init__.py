"""Shim of the timm subset the reference imports. Test infra only."""
