"""Import alias.  The package directory is named `persian-rag-system_b200/` after the reference
repository; a hyphen is not legal in a Python identifier, so `import persian_rag_system_b200`
resolves to that directory through this one-file loader."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "persian-rag-system_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
