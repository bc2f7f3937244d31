"""Generates tests/golden/rtty_fbank.npz by running the REFERENCE's own RTTY executive loop
(/root/reference/rtty.py, RTTY_Executive.run) in the build container:  python tests/golden/make_golden_rtty.py

rtty.py imports Qt, pyqtgraph, sig_proc, Plotting and profiler, none of which exist here; they are replaced by
empty stub modules (only names needed at import / class-definition time).  The loop's arithmetic — RTTY_Params,
the Kaiser window, the four quarter-symbol FFTs per symbol, 10 log10 |X|^2, flipud — is the reference's code,
unmodified, driven through a stub ring buffer; mark_bins is emptied so no decoders run, and find_sigs (which edits
the display waterfall) is replaced by a recorder of the line it is handed.  The reference tree is read, never
written, and nothing from it is copied into this repository — only the numeric outputs are saved."""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.util import rtty_input        # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/rtty.py"
FS_OUT = 48000


def load_reference():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class QWidget(object):
        pass

    stub("PyQt6")
    stub("PyQt6.QtWidgets", QWidget=QWidget)
    stub("pyqtgraph")
    stub("sig_proc", ring_buffer2=object)
    stub("Plotting", pyqtSignal=lambda *a, **k: None, pyqtSlot=lambda *a, **k: (lambda f: f))
    stub("profiler", Profiler2=object)
    spec = importlib.util.spec_from_file_location("ref_rtty", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class Driver(object):
    """The `self` handed to the reference's RTTY_Executive.run."""

    def __init__(self, mod, x):
        class P:
            pass
        self.P = P()
        self.P.FS_OUT = FS_OUT
        class Q:
            def put(self, item):
                pass
        self.q_out = Q()
        self.active = True
        self.Done = False
        self.lines = []
        self._x, self._pos = x, 0
        outer = self

        class RB:
            tag = 'RTTY'

            @property
            def nsamps(self):
                return len(outer._x) - outer._pos

            def ready(self, n):
                return True

            def pull(self, n):
                y = outer._x[outer._pos:outer._pos + n]
                outer._pos += n
                return y

        class PR:
            enabled = False
            triggered = False
        self.rb, self.pr = RB(), PR()

    def msg_handler2(self):
        pass

    def find_sigs(self, line):
        self.lines.append(np.array(line[0], dtype=np.float64))
        if len(self._x) - self._pos < self.N:
            self.Done = True


if __name__ == "__main__":
    mod = load_reference()
    mod.mark_bins = []
    R = mod.RTTY_Params(FS_OUT)
    n_sym = 12
    x = rtty_input(n_sym, R.N, FS_OUT)
    d = Driver(mod, x)
    mod.RTTY_Executive.run(d)
    lines = np.array(d.lines)
    assert lines.shape == (4 * (n_sym - 1), R.NFFT), lines.shape
    np.savez_compressed(os.path.join(HERE, "rtty_fbank.npz"), lines=lines.astype(np.float32), N=R.N, NFFT=R.NFFT,
                        NSTART=np.array(R.NSTART), NBINS=R.NBINS, frq=R.frq, n_sym=n_sym, seed=505,
                        window=np.asarray(d.window, np.float64))
    print("wrote rtty_fbank.npz", lines.shape, "N", R.N, "NFFT", R.NFFT, "NSTART", R.NSTART, "NBINS", R.NBINS)
