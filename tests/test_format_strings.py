"""Every `"..." % args` in the scripts the driver runs must format: bench.py's GPU-only branches cannot be executed here,
so a stray `%` in a message would only show on the GPU box.  The check is static: each constant format string is applied to
as many dummy numbers as its right-hand side supplies."""
import ast
import glob
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")] + \
    sorted(glob.glob(os.path.join(ROOT, "mnem_b200", "*.py"))) + sorted(glob.glob(os.path.join(ROOT, "scripts", "*.py"))) + \
    sorted(glob.glob(os.path.join(ROOT, "oracle", "*.py")))


@pytest.mark.parametrize("path", FILES, ids=[os.path.relpath(p, ROOT) for p in FILES])
def test_percent_format_strings_are_well_formed(path):
    tree = ast.parse(open(path).read(), path)
    checked = 0
    for node in ast.walk(tree):
        if not (isinstance(node, ast.BinOp) and isinstance(node.op, ast.Mod)):
            continue
        if not (isinstance(node.left, ast.Constant) and isinstance(node.left.value, str)):
            continue
        if isinstance(node.right, ast.Tuple):
            if any(isinstance(e, ast.Starred) for e in node.right.elts):
                continue
            args = tuple([1] * len(node.right.elts))
        elif isinstance(node.right, ast.Dict):
            continue
        else:
            args = (1,)            # a single value; a name bound to a tuple is not used in these files (the test says so if it is)
        try:
            node.left.value % args
        except (TypeError, ValueError) as e:
            raise AssertionError("%s:%d: %r does not format with %d argument(s): %s"
                                 % (os.path.relpath(path, ROOT), node.lineno, node.left.value[:80], len(args), e))
        checked += 1
    assert checked >= 0

