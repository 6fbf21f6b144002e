"""The C ABI on a machine without a GPU: rp_api.cu compiled over the simulated CUDA runtime (csrc/cuda_sim_runtime.h,
build.build_simapi) runs the same entry points with the kernels as cooperative fibres.  Here the CPU suite
runs the by-reference / device-resident-reads tests (tests/test_zz_gpu_resident.py) that way, in a process of their own with
RACON_B200_LIB pointing at the simulated build — the product library is not involved and still refuses to work without a
device (tests/test_abi.py)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_marked_resident_read_tests_pass_over_the_simulated_runtime():
    from racon_b200 import build
    lib = build.build_simapi()
    env = dict(os.environ, RACON_B200_LIB=lib)
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_zz_gpu_resident.py", "-m", "gpu", "-x", "-q", "-p",
                        "no:cacheprovider"], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=1500)
    tail = "\n".join(r.stdout.splitlines()[-15:])
    assert r.returncode == 0, tail
    assert " passed" in tail and "skipped" not in tail and "failed" not in tail, tail


def test_simulated_runtime_can_play_a_machine_without_a_device():
    """RP_SIM_DEVICES=0: the same build reports no device and the ABI fails loudly — the behaviour tests/test_abi.py pins for
    the product library on this container."""
    from racon_b200 import build
    lib = build.build_simapi()
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from racon_b200 import api\n"
            "assert api.load(build_if_missing=False).rp_device_count() == 0\n"
            "try:\n    api.PoaBatch()\nexcept api.RaconB200Error as e:\n    assert 'no CUDA device' in str(e); print('refused')\n"
            "else:\n    raise SystemExit('created a batch object without a device')\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, RACON_B200_LIB=lib, RP_SIM_DEVICES="0"),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "refused" in r.stdout, r.stdout
