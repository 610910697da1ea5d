"""CPU check of the integer identities behind the 8 -> 2 downmix kernel's quotient and batch metering
(tests/mix_math_check.c restates mix8_quot_raw / meter_batch4 of csrc/cmgpu_mix.cuh in plain C and compares them with
the specification on corner and random inputs, under UBSan). EXTENSION, parity unpinned like the kernel itself."""
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent


def test_mix_quotient_and_batch_meter_identities():
    exe = HERE / "_build" / "mix_math_check"
    exe.parent.mkdir(exist_ok=True)
    subprocess.run(["gcc", "-O2", "-Wall", "-Wextra", "-fsanitize=undefined", "-fno-sanitize-recover=undefined", "-o", str(exe),
                    str(HERE / "mix_math_check.c")], check=True)
    proc = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0 and proc.stdout.strip().endswith("ok"), proc.stdout[-2000:] + proc.stderr[-2000:]
