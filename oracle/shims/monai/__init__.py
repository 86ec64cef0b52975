"""Stub of the MONAI subset UNETR needs.  MONAI source is absent from /root/reference and from
this image, so there is no independent ground truth: the shim re-exports the restatement that
also ships in the product (`ucf_vit_b200/utils/unetr_blocks.py`).  Parity UNPINNED."""
