"""The XLA FFI shim (cnf_ot_b200/csrc/xla_ffi_shim.cc) compiles against a stand-in of the public FFI header
(tests/xla_ffi_standin: jaxlib is not installable here) whose binder type-checks every handler against its binding, and it
exports one handler per name cnf_ot_b200/jax_ffi.py registers."""
import ast
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "cnf_ot_b200", "csrc", "xla_ffi_shim.cc")
INC = ["-I" + os.path.join(ROOT, "tests", "xla_ffi_standin"), "-I" + os.path.join(ROOT, "include")]


def _registered_names():
  tree = ast.parse(open(os.path.join(ROOT, "cnf_ot_b200", "jax_ffi.py")).read())
  for node in tree.body:
    if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", None) == "HANDLERS":
      return [c.value for c in node.value.elts]
  raise AssertionError("HANDLERS not found")


def test_shim_compiles_and_exports_every_registered_handler(tmp_path):
  obj = str(tmp_path / "shim.o")
  r = subprocess.run(["g++", "-O0", "-std=c++17", "-c", *INC, SHIM, "-o", obj], capture_output=True, text=True)
  assert r.returncode == 0, r.stderr[-3000:]
  syms = {l.split()[-1] for l in subprocess.run(["nm", obj], capture_output=True, text=True).stdout.splitlines() if " T " in l}
  names = _registered_names()
  assert len(names) == 15 and set(names) <= syms, set(names) - syms
  # without the header the translation unit is empty (the product build never needs jaxlib)
  r = subprocess.run(["g++", "-std=c++17", "-c", "-I" + os.path.join(ROOT, "include"), SHIM, "-o", str(tmp_path / "e.o")],
                     capture_output=True, text=True)
  assert r.returncode == 0, r.stderr
  assert "Cnfot" not in subprocess.run(["nm", str(tmp_path / "e.o")], capture_output=True, text=True).stdout


def test_standin_rejects_a_handler_that_does_not_match_its_binding(tmp_path):
  src = tmp_path / "bad.cc"
  src.write_text('#include "xla/ffi/api/ffi.h"\nnamespace ffi = xla::ffi;\ntypedef struct CUstream_st* cudaStream_t;\n'
                 "static ffi::Error Impl(cudaStream_t, ffi::Buffer<ffi::F32>, float) { return ffi::Error::Success(); }\n"
                 "XLA_FFI_DEFINE_HANDLER_SYMBOL(Bad, Impl, ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()"
                 '.Arg<ffi::Buffer<ffi::F32>>().Attr<int64_t>("n").Ret<ffi::Buffer<ffi::F32>>());\n')
  r = subprocess.run(["g++", "-std=c++17", "-c", *INC, str(src), "-o", str(tmp_path / "bad.o")], capture_output=True, text=True)
  assert r.returncode != 0 and "does not match" in r.stderr
