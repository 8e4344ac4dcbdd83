"""The Go face (go/deepq) against the reference's exported surface.

No Go toolchain exists in this image (`go vet` cannot run), so the check is textual: every exported
function, method, struct field, named type and package-level variable of gold's
pkg/v1/agent/deepq/{agent,memory,policy}.go must exist in go/deepq with the same signature, and
go/deepq.Sequential must carry all eleven methods of goro's model.Model interface (model.go:17-50)
with matching signatures.  The reference side is the committed fixture
tests/golden/go_api_surface.json (tools/make_go_surface.py); when /root/reference is present the
fixture is also checked to be up to date."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_go_surface as M  # noqa: E402

FIX = json.load(open(os.path.join(ROOT, "tests/golden/go_api_surface.json")))


def test_fixture_matches_the_reference_when_it_is_present():
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "pkg/v1/agent/deepq")):
        return
    assert M.reference_surface(ref) == FIX, "run tools/make_go_surface.py"


def test_deepq_package_exports_the_reference_surface():
    ours, ref = M.repo_surface(), FIX["deepq"]
    for name, sig in ref["funcs"].items():
        assert ours["funcs"].get(name) == sig, (name, ours["funcs"].get(name), sig)
    for name, sig in ref["methods"].items():
        assert ours["methods"].get(name) == sig, (name, ours["methods"].get(name), sig)
    for name, fields in ref["structs"].items():
        assert name in ours["structs"], name
        assert ours["structs"][name][:len(fields)] == fields, (name, ours["structs"][name], fields)
    for name, sig in ref["types"].items():
        assert ours["types"].get(name) == sig, name
    assert set(ref["vars"]) <= set(ours["vars"]), set(ref["vars"]) - set(ours["vars"])


def test_sequential_implements_every_method_of_model_Model():
    ours = M.repo_surface()["methods"]
    for name, sig in FIX["model_interface"].items():
        got = ours.get(f"*Sequential.{name}")
        assert got is not None, f"Sequential lacks {name}"
        # inside the model package the interface says InputOr / *Input / Opt / ValueOr; from outside they are
        # qualified with the import alias
        want = re.sub(r"\b(InputOr|Input|Opt|ValueOr)\b", r"modelv1.\1", sig)
        assert got == want, (name, got, want)
    # the compile-time assertion a Go build would evaluate is written down too
    src = open(os.path.join(ROOT, "go/deepq/policy.go")).read()
    assert "var _ modelv1.Model = (*Sequential)(nil)" in src


def test_cgo_binding_covers_the_abi_the_go_layer_uses():
    """Every C.goldq_* symbol the cgo file calls is declared in include/goldq.h, and every method go/deepq
    calls on the handle exists in go/goldq."""
    hdr = open(os.path.join(ROOT, "include/goldq.h")).read()
    cgo = open(os.path.join(ROOT, "go/goldq/goldq.go")).read()
    for sym in set(re.findall(r"C\.(goldq_\w+)\(", cgo)):
        assert re.search(r"\b%s\s*\(" % sym, hdr), sym
    methods = set(re.findall(r"^func \(x \*Handle\) (\w+)\(", cgo, flags=re.M))
    used = set()
    for f in os.listdir(os.path.join(ROOT, "go/deepq")):
        src = open(os.path.join(ROOT, "go/deepq", f)).read()
        used |= set(re.findall(r"\b(?:s\.h|a\.handle|handle|h|to\.h)\.(\w+)\(", src))
    used -= {"Close"} - methods
    assert used <= methods, used - methods
