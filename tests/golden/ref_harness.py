"""Loads the REFERENCE's own modules (/root/reference: Tables.py, utils.py, params.py, receiver.py, Plotting.py) in the
build container so that their code can be EXECUTED — to generate golden vectors (make_golden_ref_callers.py) and to
drive a `sig_proc` implementation through the reference's real callers (tools/run_reference_callers.py).

The reference imports a dozen modules that do not exist here (Qt, pyqtgraph, SoapySDR, rtlsdr, rig_io, widgets_qt,
utilities, fileio, audio_io, xlrd, unidecode) and one that is the seam of this project (`sig_proc`).  The absent ones
are replaced by empty stub modules carrying only the names needed at import / class-definition time; `sig_proc` is
whatever module the caller passes in (the numpy oracle here, `pysdr_b200.sig_proc` on a GPU box).  The reference tree
is only read — nothing from it is written or copied into this repository; only numeric outputs are saved.

Not importable on the GPU box (no /root/reference there): GPU tests use the committed fixtures instead.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("PYSDR_REFERENCE", "/root/reference")


def available():
    return os.path.exists(os.path.join(REF_ROOT, "receiver.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Anything(object):
    """Accepts any construction / attribute / call (Qt widgets, pyqtgraph items)."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


class _RingBuffer(object):
    """Recording stand-in for dsp.ring_buffer2/3 (reference pySDR.py:103-112): keeps what was pushed."""

    def __init__(self, tag, size, PREVENT_OVERFLOW=False):
        self.tag, self.size, self.pushed, self.nsamps = tag, size, [], 0
        import queue
        self.buf = queue.Queue()

    def push(self, x):
        import numpy as np
        self.pushed.append(np.array(x, copy=True))
        self.nsamps += len(x)

    def clear(self):
        self.pushed, self.nsamps = [], 0


def _error_trap(msg='', trace=False):
    raise RuntimeError("reference error_trap: %s" % (msg,))


def load(dsp_module, names=("Tables", "utils", "params", "receiver", "Plotting"), record_rings=True):
    """Returns {name: module} of the reference's modules, executed with sig_proc = dsp_module.  record_rings: the ring
    buffers the reference builds through dsp.ring_buffer2/3 keep every pushed block (host-side plumbing, so that a run's
    player payloads can be read back)."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    facade = types.ModuleType('sig_proc')      # the seam: every name the reference takes from `sig_proc`
    facade.__dict__.update({k: v for k, v in vars(dsp_module).items() if not k.startswith('__')})
    for k, v in (('ring_buffer2', _RingBuffer), ('ring_buffer3', _RingBuffer)):
        if record_rings or k not in facade.__dict__:
            facade.__dict__[k] = v             # host-side plumbing (the numpy oracle does not restate it)
    sys.modules['sig_proc'] = facade
    qt = dict(QMessageBox=_Anything, QApplication=_Anything, QLCDNumber=_Anything, QLabel=_Anything, QWidget=object,
              QIcon=_Anything, QPixmap=_Anything, QTransform=_Anything, QFont=_Anything, Qt=_Anything())
    _stub("PyQt6")
    _stub("PyQt6.QtWidgets", **qt)
    _stub("PyQt6.QtGui", **qt)
    _stub("PyQt6.QtCore", pyqtSignal=lambda *a, **k: None, pyqtSlot=lambda *a, **k: (lambda f: f), **qt)
    _stub("widgets_qt", QTLIB="PyQt6")
    _stub("pyqtgraph", **{k: _Anything for k in ("InfiniteLine", "mkPen", "GraphicsLayoutWidget", "ImageItem", "ColorMap",
                                                 "TextItem", "PlotCurveItem", "ScatterPlotItem", "mkBrush", "mkQApp")})
    _stub("xlrd")
    _stub("unidecode", unidecode=lambda s: s)
    _stub("rig_io", bands={}, CONNECTIONS=['NONE'], RIGS=['NONE'])
    _stub("utilities", freq2band=lambda f: '', error_trap=_error_trap, whoami=lambda: '')
    _stub("rtlsdr", RtlSdr=_Anything)
    _stub("fileio")
    _stub("audio_io", AudioIO=_Anything)
    out = {}
    for name in names:
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod               # the reference's modules import each other by these names
        spec.loader.exec_module(mod)
        out[name] = mod
    return out


def run_time_params(mods, argv):
    """The reference's own RUN_TIME_PARAMS (params.py:38-472) for a command line (list of strings after the program name)."""
    old = sys.argv
    sys.argv = ["pySDR.py"] + list(argv)
    try:
        return mods["params"].RUN_TIME_PARAMS()
    finally:
        sys.argv = old
