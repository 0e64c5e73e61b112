"""Write tests/golden/reference_signatures.json from the UNMODIFIED reference.  Build container only.

    python -m oracle.make_signatures

For every top-level function of the reference's hot-path modules (laris/tools/__init__.py, laris/tools/_utils.py)
the list of positional parameters with the source text of their defaults, plus each module's ``__all__``.
tests/test_abi_and_host.py diffs the product's signatures against this list (the drop-in boundary).
"""
from __future__ import annotations

import ast
import json
import os

REF = "/root/reference/laris"
FILES = {"tools/__init__.py": "tools/__init__.py", "tools/_utils.py": "tools/_utils.py",
         "preprocessing/__init__.py": "preprocessing/__init__.py"}
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "reference_signatures.json")


def signatures(path):
    tree = ast.parse(open(path).read())
    fns, exported = {}, None
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            a = node.args
            pos = [x.arg for x in a.posonlyargs + a.args]
            defaults = [ast.unparse(d) for d in a.defaults]
            dmap = dict(zip(pos[len(pos) - len(defaults):], defaults))
            fns[node.name] = {"positional": [[p, dmap.get(p)] for p in pos],
                              "keyword_only": [x.arg for x in a.kwonlyargs]}
        elif isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id == "__all__" for t in node.targets):
            exported = [ast.literal_eval(e) for e in node.value.elts]
    return {"functions": fns, "__all__": exported}


def main():
    out = {rel: signatures(os.path.join(REF, rel)) for rel in FILES}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", OUT, {k: len(v["functions"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
