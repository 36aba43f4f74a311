"""Host logic of the stencil-run assembly path, without a GPU: the run detection, the template
programs (finite-element-solver_b200/csrc/runs_template.hpp) and the kernel's addressing scheme are
replayed by a C++ emulator (tests/host/runs_emulator.cpp) on the reference's structured meshes with
jittered coordinates (ref fem/mesh/structured.py:17-96) and compared with a plain per-pair assembly
(1e-13 relative; every covered value written exactly once, no uncovered value touched)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_run_templates_replay_on_the_host(tmp_path):
    exe = str(tmp_path / 'runs_emulator')
    src = os.path.join(ROOT, 'tests', 'host', 'runs_emulator.cpp')
    subprocess.run(['g++', '-O1', '-std=c++17', '-o', exe, src], check=True, capture_output=True, text=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'RUNS_EMULATOR PASS' in r.stdout
    # interior tets: 18 element lines, 15 pairs per node; criss-cross triangles: period 2
    assert 'covered=0.9' in r.stdout
