"""Extract the exported API surface of gold's deepq package and goro's model.Model interface from the
reference sources into tests/golden/go_api_surface.json (there is no Go toolchain here, so
tests/test_go_surface.py holds go/deepq to this fixture textually instead of with `go vet`).

    python tools/make_go_surface.py [/root/reference]
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def strip_comments(src):
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return "\n".join(re.sub(r"//.*$", "", line) for line in src.splitlines())


def norm(s):
    return re.sub(r"\s+", " ", s).strip()


def go_surface(src):
    """Exported funcs/methods (with signatures), struct types (with exported/embedded fields), named func types
    and package-level vars of one Go source file."""
    src = strip_comments(src)
    out = {"funcs": {}, "methods": {}, "structs": {}, "types": {}, "vars": []}
    for m in re.finditer(r"^func\s+(\((\w+)\s+(\*?\w+)\)\s+)?(\w+)\s*\((.*?)\)\s*(.*?)\s*\{\s*$", src, flags=re.M):
        _, _, recv, name, params, results = m.groups()
        if not name[0].isupper():
            continue
        sig = norm(f"({params}) {results}")
        if recv:
            out["methods"][f"{recv}.{name}"] = sig
        else:
            out["funcs"][name] = sig
    for m in re.finditer(r"^type\s+(\w+)\s+struct\s*\{(.*?)^\}", src, flags=re.M | re.S):
        name, body = m.groups()
        if not name[0].isupper():
            continue
        fields = []
        for line in body.splitlines():
            line = norm(line)
            if not line:
                continue
            first = line.lstrip("*").split(".")[-1] if " " not in line else line.split()[0]
            if first[0].isupper():
                fields.append(line)
        out["structs"][name] = fields
    for m in re.finditer(r"^type\s+(\w+)\s+(func\(.*)$", src, flags=re.M):
        if m.group(1)[0].isupper():
            out["types"][m.group(1)] = norm(m.group(2))
    for m in re.finditer(r"^var\s+(\w+)\s*=", src, flags=re.M):
        if m.group(1)[0].isupper():
            out["vars"].append(m.group(1))
    return out


def interface_methods(src, name):
    src = strip_comments(src)
    m = re.search(r"^type\s+%s\s+interface\s*\{(.*?)^\}" % name, src, flags=re.M | re.S)
    out = {}
    for line in m.group(1).splitlines():
        line = norm(line)
        mm = re.match(r"(\w+)\s*\((.*?)\)\s*(.*)$", line)
        if mm:
            out[mm.group(1)] = norm(f"({mm.group(2)}) {mm.group(3)}")
    return out


def merge(parts):
    out = {"funcs": {}, "methods": {}, "structs": {}, "types": {}, "vars": []}
    for p in parts:
        for k in ("funcs", "methods", "structs", "types"):
            out[k].update(p[k])
        out["vars"] += p["vars"]
    out["vars"] = sorted(set(out["vars"]))
    return out


def reference_surface(ref):
    d = os.path.join(ref, "pkg/v1/agent/deepq")
    deepq = merge([go_surface(open(os.path.join(d, f)).read()) for f in ("agent.go", "memory.go", "policy.go")])
    model = open(os.path.join(ref, "vendor/github.com/aunum/goro/pkg/v1/model/model.go")).read()
    return {"source": "aunum/gold pkg/v1/agent/deepq/{agent,memory,policy}.go + goro pkg/v1/model/model.go:17-50",
            "deepq": deepq, "model_interface": interface_methods(model, "Model")}


def repo_surface():
    d = os.path.join(ROOT, "go/deepq")
    return merge([go_surface(open(os.path.join(d, f)).read()) for f in sorted(os.listdir(d)) if f.endswith(".go")])


if __name__ == "__main__":
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = os.path.join(ROOT, "tests/golden/go_api_surface.json")
    json.dump(reference_surface(ref), open(out, "w"), indent=1, sort_keys=True)
    print("wrote", out)
