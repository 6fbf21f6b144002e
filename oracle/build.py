"""Builds the checkers (TEST INFRASTRUCTURE): oracle/_build/libpoa_oracle.so and, where /root/reference
exists, oracle/_ref/libracon_ref.so — both via oracle/Makefile."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def _make(target):
    r = subprocess.run(["make", "-C", HERE, target], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed (%s):\n%s" % (target, r.stdout))


def build_oracle():
    _make("oracle")
    return os.path.join(HERE, "_build", "libpoa_oracle.so")


def build_ref():
    out = os.path.join(HERE, "_ref", "libracon_ref.so")
    if os.path.isdir("/root/reference/src"):
        _make("ref")
    return out if os.path.exists(out) else None
