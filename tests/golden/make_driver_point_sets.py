"""Fingerprints of the point sets the REFERENCE'S OWN driver scripts build, for tests/test_preprocess.py.

    PYTHONPATH=/root/repo python tests/golden/make_driver_point_sets.py          -> tests/golden/driver_point_sets.npz

The four scripts define their training sets in the `if __name__ == "__main__":` body (plate:871-929, inf:634-705, semi:667-739,
conf:881-968), which cannot be imported.  This generator reads each script from /root/reference, takes the lines of that body up to the
first `with tf.device` (the model construction) and executes them, as written, in the namespace of the imported module (its own
GenDistPt / DelSrcPT / CartGrid / ... helpers), with the numpy global stream seeded like the scripts do (np.random.seed(1111), plate:22).
Stand-ins: matplotlib is stubbed (oracle/tf1_shim.py), and pyDOE -- absent here -- is replaced by the package's restatement of its
published _lhsclassic algorithm (pinn_elastodynamics_b200.preprocess.lhs), so the fixture pins everything EXCEPT the LHS draw itself.
Stored per array: shape, sum, sum of squares, and a SHA-256 of the float64 bytes.  Nothing at test time reads /root/reference."""
import hashlib
import os
import sys
import textwrap

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import tf1_shim as tf                              # noqa: E402
from pinn_elastodynamics_b200 import preprocess as P           # noqa: E402

REF = '/root/reference'
SCRIPTS = {'plate': ('PlateHoleQuarter/train/train.py', dict(Collo='XYT_c', HOLE='HOLE', IC='IC', LF='LF', RT='RT', UP='UP', LW='LW', DIST='DIST')),
           'semi': ('ElasticWaveSemiInfinite/ElasticWave.py', dict(Collo='XYT_c', SRC='SRC', IC='IC', UP='UP')),
           'inf': ('ElasticWaveInfinite/ElasticWave.py', dict(Collo='XYT_c', SRC='SRC', IC='IC', UP='UP')),
           'conf': ('ElasticWaveConfined/ElasticWave.py', dict(Collo='XYT_c', SRC='SRC', IC='IC', FIXED='FIXED', DIST='DIST'))}


def fingerprint(a):
    a = np.ascontiguousarray(np.asarray(a, np.float64))
    return np.array(a.shape, np.int64), np.array([a.sum(), (a * a).sum()]), np.frombuffer(hashlib.sha256(a.tobytes()).digest(), np.uint8)


class _Plot:
    """matplotlib stand-in for the driver bodies: every attribute / call returns another stand-in, `plt.subplots()` unpacks into two"""
    def __getattr__(self, name):
        return _Plot()

    def __call__(self, *a, **k):
        return _Plot()

    def __iter__(self):
        return iter((_Plot(), _Plot()))


def driver_body(path):
    lines = open(path).read().split('\n')
    i0 = next(i for i, l in enumerate(lines) if l.startswith('if __name__ =='))
    i1 = next(i for i in range(i0, len(lines)) if 'with tf.device' in lines[i])
    return textwrap.dedent('\n'.join(lines[i0 + 1:i1]))


if __name__ == '__main__':
    OUT = {}
    for kind, (rel, names) in SCRIPTS.items():
        mod = tf.load_reference_module(os.path.join(REF, rel), 'ref_driver_' + kind)
        ns = dict(mod.__dict__)
        ns['lhs'] = P.lhs
        ns['plt'] = _Plot()
        np.random.seed(1111)
        exec(compile(driver_body(os.path.join(REF, rel)), rel + ':__main__', 'exec'), ns)
        for ours, theirs in names.items():
            sh, sm, dg = fingerprint(ns[theirs])
            OUT[f'{kind}_{ours}_shape'], OUT[f'{kind}_{ours}_sums'], OUT[f'{kind}_{ours}_sha256'] = sh, sm, dg
            print(kind, ours, tuple(sh), sm)
    out = os.path.join(HERE, 'driver_point_sets.npz')
    np.savez_compressed(out, **OUT)
    print('wrote', out, os.path.getsize(out), 'bytes')
