"""Writes tests/golden/randomvariable_interface.json: the method names of the reference's RandomVariable interface
(src/main/java/net/finmath/stochastic/RandomVariable.java), abstract and default, read from the reference checkout in this container.
The names travel as a fixture because /root/reference does not exist where the tests run later.

    python tests/golden/make_interface_list.py [/root/reference]
"""
import json
import os
import re
import sys

root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = open(os.path.join(root, "src/main/java/net/finmath/stochastic/RandomVariable.java")).read()
names = set()
for m in re.finditer(r"^\s*(?:default\s+)?(?:[\w<>\[\],\s\.]+?)\s+(\w+)\s*\(([^)]*)\)\s*(?:;|\{)", src, re.M):
    if m.group(1) not in ("if", "for", "while", "switch", "return", "new"):
        names.add(m.group(1))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "randomvariable_interface.json")
json.dump({"source": "net/finmath/stochastic/RandomVariable.java", "methods": sorted(names)}, open(out, "w"), indent=1)
print(len(names), "method names ->", out)
