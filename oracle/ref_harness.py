"""Import the UNMODIFIED reference transcoder modules (test infrastructure only).

Only usable where ``/root/reference`` is mounted (the build container); the GPU
box has no such path, so nothing under ``tests -m gpu``, ``smoke()`` or
``bench.py`` may call this.  It exists to (a) validate ``oracle/scorer.py`` and
``oracle/tables.py`` against the real reference code and (b) generate the
committed fixtures under ``tests/golden/`` (see ``oracle/make_golden.py``).

What can be imported: screen.py, colours.py, palette.py, video.py, opcodes.py,
machine.py, symbol_table.py, video_mode.py, frame_grabber.py -- given
  * ``np.bool8`` alias (reference screen.py:42 uses the numpy<2 name),
  * stub ``colormath.color_objects`` (palette.py:6-15), ``skvideo.io``
    (frame_grabber.py:10), ``audioread`` and ``librosa`` (audio.py:5-6, only imported)
    from ``oracle/ref_shims``,
  * CWD containing ``player/iivision.dbg`` while opcodes.py is imported
    (opcodes.py:173 opens it by relative path at import time).
What cannot: make_data_tables.py (needs colormath 3.0.0, weighted-levenshtein
0.2.2 and etaprogress, none installed, none vendored).
"""

import contextlib
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("IIV_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")
_MODULES = (
    "colours", "palette", "video_mode", "symbol_table", "machine", "opcodes",
    "screen", "frame_grabber", "video", "audio", "movie",
)
_cache = None


def available() -> bool:
    return os.path.isfile(
        os.path.join(REFERENCE_ROOT, "transcoder", "screen.py"))


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def load() -> types.SimpleNamespace:
    """Returns a namespace whose attributes are the reference's own modules."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_ROOT)

    import numpy as np
    if not hasattr(np, "bool8"):
        np.bool8 = np.bool_

    tdir = os.path.join(REFERENCE_ROOT, "transcoder")
    for p in (tdir, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    ns = types.SimpleNamespace()
    with _cwd(REFERENCE_ROOT):
        for name in _MODULES:
            mod = importlib.import_module(name)
            if not os.path.abspath(mod.__file__).startswith(tdir):
                raise RuntimeError(
                    "module %r resolved to %s, not the reference" % (
                        name, mod.__file__))
            setattr(ns, name, mod)
    # video.py:90 prints the mean priority on every encode_frame call.
    ns.video.print = lambda *a, **k: None
    _cache = ns
    return ns


def install_tables(ns, mode_name: str, tables_by_palette_id: dict) -> None:
    """Feeds symmetric uint16[(n_offsets, 4**bits)] tables to the reference
    scorer in place of the .npz loader (screen.py:343-367)."""
    cls = {"HGR": ns.screen.HGRBitmap, "DHGR": ns.screen.DHGRBitmap}[mode_name]

    def _edit_distances(klass, palette_id, _t=tables_by_palette_id):
        return _t[palette_id.value]

    cls.edit_distances = classmethod(_edit_distances)
