"""Import shim: the package directory is ``openmp-blas_b200/`` (a hyphen is not importable),
so ``import openmp_blas_b200`` loads it from there under this name."""
import importlib.util as _ilu
import pathlib as _pl
import sys as _sys

_pkg = _pl.Path(__file__).resolve().with_name("openmp-blas_b200")
_spec = _ilu.spec_from_file_location(__name__, _pkg / "__init__.py",
                                     submodule_search_locations=[str(_pkg)])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
