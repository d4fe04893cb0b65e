"""Batched device engines: thin torch-tensor owners around the C ABI."""
