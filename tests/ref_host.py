"""Run pieces of the reference's HOST code (cuda_lib_gl.py, Python 2) under Python 3 -- TEST INFRASTRUCTURE.

The reference cannot be imported (PyCUDA, OpenGL, print statements), but several of its methods are plain NumPy
and already valid Python 3 apart from `xrange`.  Their source TEXT is read from /root/reference at test time
(never copied into the repo), dedented and exec'd; `self` is a mock object prepared by the test."""
import os
import textwrap
import time

import numpy as np

REF = "/root/reference"                      # fixed: the reference tree of this container, never redirected by the environment
PATH = os.path.join(REF, "cuda_lib_gl.py")


def available():
    """The reference tree is present and its execution has not been switched off (GRAAL_RUN_REFERENCE=0: rely on the
    frozen golden vectors only)."""
    return os.path.exists(PATH) and os.environ.get("GRAAL_RUN_REFERENCE", "1") != "0"


class _NumpyOfItsTime:
    """numpy as the reference imported it (1.x): `np.lib.arraysetops.*` was public."""
    lib = __import__("types").SimpleNamespace(arraysetops=__import__("types").SimpleNamespace(setdiff1d=np.setdiff1d, intersect1d=np.intersect1d))

    def __getattr__(self, name):
        return getattr(np, name)


NP = _NumpyOfItsTime()


def _lines():
    return open(PATH).read().split("\n")


def method(name, extra_ns=None, py2=False):
    """The reference method `name` of class sampler as a Python 3 function f(self, ...).  ``py2``: translate the Python 2
    idioms of the text mechanically (print statements -> pass, the integer division `omega_f / self.n_modif_metropolis`
    -> //); nothing else is touched."""
    lines = _lines()
    start = next(i for i, l in enumerate(lines) if l.startswith("    def %s(" % name))
    end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith("    def ") or (lines[i] and not lines[i].startswith(" ")))
    body = lines[start:end]
    if py2:
        out = []
        for l in body:
            st = l.lstrip()
            if st.startswith("print ") or st == "print":
                l = l[:len(l) - len(st)] + "pass"
            out.append(l.replace("omega_f / self.n_modif_metropolis", "omega_f // self.n_modif_metropolis"))
        body = out
    src = textwrap.dedent("\n".join(body))
    from scipy import stats
    ns = {"np": NP, "xrange": range, "time": time, "stats": stats}
    ns.update(extra_ns or {})
    exec(compile(src, "%s:%s" % (PATH, name), "exec"), ns)
    return ns[name]


def block(after_def, first_marker, last_marker):
    """The statements of method `after_def` from the first line containing `first_marker` to the first following
    line containing `last_marker` (inclusive), dedented: returned as a code object to exec in a namespace."""
    lines = _lines()
    d = next(i for i, l in enumerate(lines) if l.startswith("    def %s(" % after_def))
    a = next(i for i in range(d, len(lines)) if first_marker in lines[i])
    b = next(i for i in range(a, len(lines)) if last_marker in lines[i])
    src = textwrap.dedent("\n".join(lines[a:b + 1]))
    return compile(src, "%s:%s[%d:%d]" % (PATH, after_def, a + 1, b + 1), "exec")


LOADER = os.path.join(REF, "simulation_loader.py")


def loader_method(name, extra_ns=None):
    """Method `name` of class simulation (simulation_loader.py) as a Python 3 function: the Python 2 print STATEMENTS of
    the text are replaced by `pass` (same indentation), nothing else is touched."""
    lines = open(LOADER).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("    def %s(" % name))
    end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith("    def ") or (lines[i] and not lines[i].startswith(" ")))
    body = []
    for l in lines[start:end]:
        st = l.lstrip()
        body.append(l[:len(l) - len(st)] + "pass" if st.startswith("print ") or st == "print" else l)
    src = textwrap.dedent("\n".join(body))
    ns = {"np": NP, "xrange": range, "time": time}
    ns.update(extra_ns or {})
    exec(compile(src, "%s:%s" % (LOADER, name), "exec"), ns)
    return ns[name]


PYRAMID = os.path.join(REF, "pyramid_sparse.py")


def pyramid_function(name, extra_ns=None):
    """Module-level function `name` of pyramid_sparse.py as a Python 3 function.  Python 2 idioms of the text are
    translated mechanically, nothing else is touched: print statements -> pass, d.has_key(k) -> (k in d),
    `x = d.keys()` -> `x = list(d.keys())` (the text sorts such lists in place)."""
    import re
    lines = open(PYRAMID).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("def %s(" % name))
    end = next(i for i in range(start + 1, len(lines)) if lines[i] and not lines[i][0] in " #")
    body = []
    for l in lines[start:end]:
        st = l.lstrip()
        if st.startswith("print ") or st == "print":
            l = l[:len(l) - len(st)] + "pass"
        l = re.sub(r"(\w+(?:\[[^\]]+\])*)\.has_key\(([^)]+)\)", r"(\2 in \1)", l)
        l = re.sub(r"=\s*(\w+)\.keys\(\)\s*$", r"= list(\1.keys())", l)
        body.append(l)
    ns = {"np": NP, "xrange": range, "time": time}
    ns.update(extra_ns or {})
    exec(compile("\n".join(body), "%s:%s" % (PYRAMID, name), "exec"), ns)
    return ns[name]
