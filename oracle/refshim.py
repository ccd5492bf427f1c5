"""SymPy/Python-3 compatibility shim for importing the *reference* OpenSBLI front end.

TEST INFRASTRUCTURE ONLY.  The reference (Python-2.7 / SymPy-1.1 era) does not import under
Python 3.12 / SymPy 1.14 without h5py / matplotlib.  This module installs the small set of
module aliases and monkey-patches documented in SURVEY.md Appendix D so that the *unmodified*
reference sources under /root/reference can be imported and run in this container.  Nothing
in /root/reference is modified.  The product package (`opensbli_b200`) re-uses `install()`
only when an app is driven through the reference front end; it is not on the compute path.

Usage:   import refshim; refshim.install()      # before and after `import opensbli`
"""
import sys
import types
import tokenize
import itertools

_DONE = {'pre': False, 'post': False}


def install_pre():
    """Module aliases that must exist before `import opensbli`."""
    if _DONE['pre']:
        return
    import sympy
    import sympy.printing.c as _c
    sys.modules.setdefault('sympy.printing.ccode', _c)                    # opsc.py:8
    compat = types.ModuleType('sympy.core.compatibility')
    from sympy.utilities.iterables import is_sequence
    compat.is_sequence = is_sequence
    compat.exec_ = lambda code, g=None, l=None: exec(code, g, l)          # parsing.py:398-399
    sys.modules['sympy.core.compatibility'] = compat
    tok = types.ModuleType('sympy.parsing.sympy_tokenize')
    tok.NAME, tok.OP = tokenize.NAME, tokenize.OP
    sys.modules['sympy.parsing.sympy_tokenize'] = tok                      # parsing.py:11
    if 'h5py' not in sys.modules:
        try:
            import h5py  # noqa: F401
        except Exception:
            sys.modules['h5py'] = types.ModuleType('h5py')                 # helperfunctions.py:8
    from sympy import Symbol, Function
    Symbol.__call__ = lambda self, *a: Function(self.name)(*a)             # SymPy-1.1 semantics
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl, plt = types.ModuleType('matplotlib'), types.ModuleType('matplotlib.pyplot')
        plt.style = types.SimpleNamespace(use=lambda *a, **k: None)
        mpl.pyplot = plt
        sys.modules.setdefault('matplotlib', mpl)
        sys.modules.setdefault('matplotlib.pyplot', plt)                   # katzer_init.py:6,16
    _DONE['pre'] = True


def install_post():
    """Monkey patches applied after `import opensbli`."""
    if _DONE['post']:
        return
    import opensbli  # noqa: F401
    from opensbli.code_generation.latex import LatexWriter
    LatexWriter.write_expression = lambda self, *a, **k: None             # latex.py:150
    LatexWriter.latexify_expression = lambda self, *a, **k: ''            # latex.py:79
    from sympy import flatten, S, Integer, factor, Equality
    from sympy.tensor import IndexedBase, Indexed
    import opensbli.core.opensblifunctions as F

    def _struct(self, n=None):                                             # opensblifunctions.py:20-31,50-62
        idx = flatten([p.get_indices() for p in self.args if p.get_indices])
        ib = IndexedBase(self.__class__.__name__)
        out = ib[tuple(idx)] if n else ib[idx]
        out.expression = self
        try:
            out.is_commutative = False
        except AttributeError:
            pass
        return out
    F.KD.structure = lambda self: _struct(self)
    F.LC.structure = lambda self: _struct(self, 3)
    import opensbli.schemes.spatial.shock_capturing as SC
    from opensbli.equation_types.opensbliequations import OpenSBLIEq, OpenSBLIEquation

    def _gen(self, lhs, rhs):                                              # shock_capturing.py:233-245
        assert lhs.shape == rhs.shape
        return [OpenSBLIEq(v, factor(rhs[i])) if rhs[i] != 0 else S.Zero for i, v in enumerate(lhs)]
    SC.EigenSystem.generate_equations_from_matrices = _gen
    from opensbli.code_generation.opsc import OPSCCodePrinter as P
    P._print_GroupedPiecewise = P._print_Piecewise                         # teno.py:456-459
    import opensbli.core.opensbliobjects as O

    def _gnew(cls, label, number, **kw):                                   # opensbliobjects.py:503-508
        r = Indexed.__new__(cls, str(label), number, **kw)
        r.number = number
        r._args = (r.base, Integer(number))
        return r
    O.Grididx.__new__ = staticmethod(_gnew)
    import opensbli.equation_types.metric as M

    def _sd(cls):                                                          # metric.py:207-212
        eq = []
        for ijk in itertools.product(range(cls.ndim), repeat=3):
            x, y = cls.SD_metrics[ijk], cls.SD_evaluations[ijk]
            if x != 0:
                e = OpenSBLIEq(x, y)
                if isinstance(e, Equality):
                    eq.append(OpenSBLIEquation(e.lhs, e.rhs))
        cls.sdequations = eq
    M.MetricsEquation.generate_sd_metrics_equations = _sd
    _DONE['post'] = True


def install(reference_root='/root/reference'):
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    install_pre()
    install_post()
