"""Import shim: the product package lives in ``radiant-rag_b200/`` (the name the build
contract fixes); a hyphen is not importable, so this package points its ``__path__`` at
that directory and executes its ``__init__``."""
from pathlib import Path as _Path

_real = _Path(__file__).resolve().parent.parent / "radiant-rag_b200"
__path__ = [str(_real)]
__file__ = str(_real / "__init__.py")
exec(compile((_real / "__init__.py").read_text(), __file__, "exec"))
