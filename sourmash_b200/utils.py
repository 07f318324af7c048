"""FFI calling convention helpers (the reference's counterpart: src/sourmash/utils.py:9-78)."""
from ._lowlevel import ffi, lib
from .exceptions import _exception_for


def decode_str(s, free=True):
    """SourmashStr -> python str (frees owned strings)."""
    try:
        if s.data == ffi.NULL or s.len == 0:
            return ""
        return ffi.unpack(s.data, s.len).decode("utf-8", "replace")
    finally:
        if free and s.owned:
            p = ffi.new("SourmashStr *", s)
            lib.sourmash_str_free(p)


def rustcall(func, *args):
    """Clear the thread-local error slot, call, poll the error code, raise if set."""
    lib.sourmash_err_clear()
    rv = func(*args)
    code = lib.sourmash_err_get_last_code()
    if not code:
        return rv
    msg = decode_str(lib.sourmash_err_get_last_message())
    lib.sourmash_err_clear()
    raise _exception_for(code)(msg)


class RustObject:
    """Owner of an opaque C handle; same protocol as the reference's RustObject."""
    __dealloc_func__ = None
    _objptr = None
    _shared = False

    def __init__(self):
        raise TypeError("Cannot instantiate %r objects" % self.__class__.__name__)

    @classmethod
    def _from_objptr(cls, ptr, shared=False):
        rv = object.__new__(cls)
        rv._objptr = ptr
        rv._shared = shared
        return rv

    def _get_objptr(self):
        if not self._objptr:
            raise RuntimeError("Object is closed")
        return self._objptr

    def _methodcall(self, func, *args):
        return rustcall(func, self._get_objptr(), *args)

    def __del__(self):
        if self._objptr is None or self._shared:
            return
        f = self.__class__.__dealloc_func__
        if f is not None and lib is not None:
            try:
                f(self._objptr)
            finally:
                self._objptr = None
