"""Load the UNMODIFIED reference modules from /root/reference/code on CPU.

Used ONLY by tests/golden/make_golden.py (fixture generation, in the build
container).  Nothing here is imported by the product, the gpu tests, smoke()
or bench.py: /root/reference does not exist on the GPU box.

Two mechanical shims:
  1. stub modules for imports that are not installed and not on the hot path
     (line_profiler, ipdb, pyxis, matplotlib);
  2. textual "cuda" -> "cpu" substitution of the hard-coded device strings
     (utils.py:19, load.py:25, models.py:19,29,231,353).
No reference source is copied into this repository; it is read and exec'd.
"""
import sys
import types

REF = "/root/reference/code"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_reference():
    class _LP:
        def print_stats(self, *a, **k):
            pass

        def __call__(self, f):
            return f

    _stub("line_profiler", LineProfiler=_LP)
    _stub("ipdb")
    _stub("pyxis")
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    mods = {}
    for name in ("constants", "utils", "load", "models"):
        src = open(f"{REF}/{name}.py").read().replace('"cuda"', '"cpu"')
        m = types.ModuleType(name)
        m.__file__ = f"{REF}/{name}.py"
        sys.modules[name] = m
        exec(compile(src, m.__file__, "exec"), m.__dict__)
        mods[name] = m
    return mods
