"""Read-only ZipStorage over the library's zip reader.

Mirrors the reference's ``ZipStorage`` (src/sourmash/sbt_storage.py:93-200, backed by
src/core/src/ffi/storage.rs:15-141): ``load`` a member, list ``_filenames`` / ``list_sbts``,
``subdir`` lookup.  Writing zip files stays with Python's ``zipfile`` in the reference
(``_RwZipStorage``); it is not on any GPU path and is not provided here.
"""
import os

from ._ffi import RustObject, decode_str, rustcall
from ._lowlevel import ffi, lib


class ZipStorage(RustObject):
    __dealloc_func__ = lib.zipstorage_free

    def __init__(self, path, *, mode="r"):
        if mode != "r":
            raise NotImplementedError("ZipStorage is read-only here; write collections with zipfile")
        path = os.path.abspath(str(path))
        raw = path.encode("utf-8")
        self._objptr = rustcall(lib.zipstorage_new, raw, len(raw))
        self._shared = False

    @staticmethod
    def can_open(location):
        "True when `location` is a zip file (sbt_storage.py:96-98 uses zipfile.is_zipfile)."
        try:
            with open(location, "rb") as fh:
                return fh.read(4) in (b"PK\x03\x04", b"PK\x05\x06")
        except OSError:
            return False

    @property
    def path(self):
        return decode_str(self._methodcall(lib.zipstorage_path))

    @property
    def subdir(self):
        return decode_str(self._methodcall(lib.zipstorage_subdir))

    @subdir.setter
    def subdir(self, value):
        raw = value.encode("utf-8")
        self._methodcall(lib.zipstorage_set_subdir, raw, len(raw))

    def _string_list(self, func):
        size = ffi.new("uintptr_t *")
        arr = self._methodcall(func, size)
        return [decode_str(arr[i][0]) for i in range(size[0])]

    def _filenames(self):
        "Member names in central-directory order (zipfile.infolist() order)."
        return self._string_list(lib.zipstorage_filenames)

    def list_sbts(self):
        return self._string_list(lib.zipstorage_list_sbts)

    def load(self, path):
        "Bytes of one member; a missing member raises FileNotFoundError (sbt_storage.py:152-170)."
        raw = path.encode("utf-8")
        size = ffi.new("uintptr_t *")
        try:
            buf = self._methodcall(lib.zipstorage_load, raw, len(raw), size)
        except ValueError:
            raise FileNotFoundError(path)
        try:
            return bytes(ffi.buffer(buf, size[0]))
        finally:
            lib.nodegraph_buffer_free(ffi.cast("uint8_t *", buf), size[0])

    def save(self, path, content, *, overwrite=False, compress=False):
        raise NotImplementedError

    def close(self):
        pass

    def flush(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
