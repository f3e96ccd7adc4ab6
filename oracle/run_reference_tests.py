"""
ORACLE / TEST INFRASTRUCTURE ONLY.

Pins the greensim shim (oracle/greensim) by running the reference's *own* hot-path tests, unmodified, from
/root/reference/tests against the reference's *unmodified* itsim package on top of the shim.  This is the
"check the oracle against every known-answer test the reference holds for this path" step (SURVEY.md section 4
star rows, section 8-c).  Only usable where /root/reference exists (the build container, not the GPU box).

Out of scope (need flask / warlock / requests, none of which touch the simulation path): tests/datastore,
tests/integ/test_int_datastore.py.

Each test file runs in its own interpreter: tests/network/service/dhcp/test_dhcp_server.py:50-53 starts
``patch("itsim.network.packet.Packet")`` in a default argument (i.e. at collection time) and never stops it, which
poisons any test module collected after it in the same process (a collection-order artefact of the reference's
suite, not of the engine under it).

Usage:  python oracle/run_reference_tests.py [--file REL_PATH ...] [extra pytest args]
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("ITSIM_REFERENCE", "/root/reference")


EXCLUDED = ("tests/datastore/", "tests/integ/test_int_datastore.py")


def hot_path_test_files():
    found = []
    for root, _, files in os.walk(os.path.join(REFERENCE, "tests")):
        for name in sorted(files):
            if name.startswith("test_") and name.endswith(".py"):
                rel = os.path.relpath(os.path.join(root, name), REFERENCE)
                if not rel.startswith(EXCLUDED):
                    found.append(rel)
    return sorted(found)


def run_one(rel_path, extra=()) -> int:
    """Runs one reference test file in this interpreter (call from a fresh process)."""
    import pytest
    sys.path.insert(0, REFERENCE)
    sys.path.insert(0, HERE)  # makes ``import greensim`` resolve to the shim
    sys.dont_write_bytecode = True  # /root/reference is read-only
    return int(pytest.main([os.path.join(REFERENCE, rel_path), "--rootdir", "/tmp", "-p", "no:cacheprovider", "-q",
                            *extra]))


def main(argv=None) -> int:
    import subprocess
    argv = list(sys.argv[1:] if argv is None else argv)
    if not os.path.isdir(os.path.join(REFERENCE, "itsim")):
        print(f"reference not found at {REFERENCE}", file=sys.stderr)
        return 2
    if argv and argv[0] == "--one":
        return run_one(argv[1], argv[2:])
    failures = 0
    for rel in hot_path_test_files():
        proc = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", rel, *argv],
                              capture_output=True, text=True)
        tail = (proc.stdout.strip().splitlines() or ["<no output>"])[-1]
        print(f"{'ok  ' if proc.returncode == 0 else 'FAIL'} {rel}: {tail}")
        if proc.returncode != 0:
            failures += 1
            print(proc.stdout[-3000:])
            print(proc.stderr[-2000:])
    print(f"reference hot-path test files failing on the shim: {failures}")
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
