"""No linter is installed in this image, so a small scope-aware check stands in for pyflakes' "undefined name": every name a
function / class body resolves as a GLOBAL must be defined at module level (assigned, imported, def / class) or be a
builtin.  Catches the classic late edit that uses ``nn.`` in a module that only imports ``torch``."""

import builtins
import symtable
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
FILES = sorted([*(ROOT / "cinema_b200").rglob("*.py"), ROOT / "bench.py", ROOT / "__graft_entry__.py", *(ROOT / "oracle").glob("*.py"),
                *(ROOT / "tools").glob("*.py"), *(ROOT / "tests").glob("*.py"), ROOT / "tests" / "golden" / "make_golden.py"])
MODULE_ATTRS = {"__file__", "__name__", "__doc__", "__spec__", "__package__", "__builtins__", "__path__", "__class__"}


def undefined_globals(path: Path) -> list[tuple[str, str]]:
    top = symtable.symtable(path.read_text(), str(path), "exec")
    defined = {s.get_name() for s in top.get_symbols() if s.is_assigned() or s.is_imported() or s.is_namespace()}
    bad: list[tuple[str, str]] = []

    def walk(tab: symtable.SymbolTable) -> None:
        for s in tab.get_symbols():
            name = s.get_name()
            if not s.is_referenced() or name in defined or hasattr(builtins, name) or name in MODULE_ATTRS:
                continue
            if tab is top:
                if not (s.is_assigned() or s.is_imported() or s.is_namespace()):
                    bad.append(("<module>", name))
            elif s.is_global() and not s.is_declared_global():
                bad.append((tab.get_name(), name))
            elif s.is_declared_global() and name not in defined:
                bad.append((tab.get_name(), name))
        for child in tab.get_children():
            walk(child)

    walk(top)
    return bad


@pytest.mark.parametrize("path", FILES, ids=lambda p: str(p.relative_to(ROOT)))
def test_no_undefined_global_names(path):
    assert undefined_globals(path) == []
