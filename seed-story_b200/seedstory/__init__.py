"""seedstory_b200 host package: ctypes C-ABI binding, op front end and the engines behind src.* ."""
from . import _capi  # noqa: F401
