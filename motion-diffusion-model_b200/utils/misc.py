"""Attribute forwarding for model wrappers (reference utils/misc.py:19-35)."""
import torch


def wrapped_getattr(self, name, default=None, wrapped_member_name="model"):
    """Look `name` up on the wrapper first and on the wrapped model second.  For nn.Module wrappers the
    Module.__getattr__ lookup (parameters / buffers / sub-modules) has to run first to avoid infinite recursion."""
    if isinstance(self, torch.nn.Module):
        try:
            return torch.nn.Module.__getattr__(self, name)
        except AttributeError:
            inner = torch.nn.Module.__getattr__(self, wrapped_member_name)
            return getattr(inner, name, default)
    return getattr(getattr(self, wrapped_member_name), name, default)
