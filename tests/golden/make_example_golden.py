"""Generate tests/golden/examples/*.out by RUNNING THE REFERENCE HERE:
the unmodified example / unit-test sources, compiled by path from
/root/reference against the reference's own nvector_serial
(oracle/_ref/bin/<prog>_serial, built by tests/c/Makefile).

Where the reference ships an expected output next to the example (the files its
own test runner diffs against), the freshly generated output is also compared
with it and the result recorded in tests/golden/examples/MANIFEST.json.

    make -C oracle ref && make -C tests/c && python tests/golden/make_example_golden.py
"""
import json
import subprocess
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
BIN = ROOT / "oracle" / "_ref" / "bin"
OUT = Path(__file__).resolve().parent / "examples"
REF = Path("/root/reference")

# program -> (args, shipped expected output inside the reference tree or None)
PROGRAMS = {
    "cvDiurnal_kry": ([], "examples/cvode/serial/cvDiurnal_kry.out"),
    "ark_heat1D": ([], "examples/arkode/C_serial/ark_heat1D.out"),
    "ark_heat2D": ([], "examples/arkode/CXX_serial/ark_heat2D.out"),
    "idaHeat2D_kry": ([], "examples/ida/serial/idaHeat2D_kry.out"),
    "kinFoodWeb_kry": ([], "examples/kinsol/serial/kinFoodWeb_kry.out"),
    "kinLaplace_picard_kry": ([], "examples/kinsol/serial/kinLaplace_picard_kry.out"),
    "test_sunlinsol_spgmr": (None, None),   # several argument sets, see SUNLS
    "test_sunlinsol_spfgmr": (None, None),
    "test_sunlinsol_pcg": (None, None),
}
# CTest argument sets of the reference (test/unit_tests/sunlinsol/*/serial/CMakeLists.txt)
SUNLS = {
    "test_sunlinsol_spgmr": [["100", "1", "1", "100", "1e-13", "0"], ["100", "2", "1", "100", "1e-13", "0"],
                             ["100", "1", "2", "100", "1e-13", "0"], ["100", "2", "2", "100", "1e-13", "0"]],
    "test_sunlinsol_spfgmr": [["100", "1", "100", "1e-13", "0"], ["100", "2", "100", "1e-13", "0"]],
    "test_sunlinsol_pcg": [["100", "500", "1e-13", "0"]],
}


def main():
    OUT.mkdir(exist_ok=True)
    manifest = {}
    for prog, (args, shipped) in PROGRAMS.items():
        argsets = SUNLS[prog] if args is None else [args]
        for a in argsets:
            tag = prog + ("_" + "_".join(a) if a else "")
            t0 = time.time()
            r = subprocess.run([str(BIN / f"{prog}_serial"), *a], capture_output=True, text=True)
            dt = time.time() - t0
            (OUT / f"{tag}.out").write_text(r.stdout)
            entry = {"program": prog, "args": a, "returncode": r.returncode, "serial_seconds": round(dt, 3)}
            if shipped and (REF / shipped).exists():
                entry["identical_to_reference_shipped_out"] = (REF / shipped).read_text() == r.stdout
                entry["reference_shipped_out"] = shipped
            manifest[tag] = entry
            print(tag, entry)
    (OUT / "MANIFEST.json").write_text(json.dumps(manifest, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()
