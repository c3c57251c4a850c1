"""Static SASS size of step_warp<11,2> for a set of extra nvcc flags (CPU only: nvcc + cuobjdump):
    python tools/studies/sass_count.py [-DFLAG ...]
Prints registers / spills, the instruction count of the kernel, of its interior-point loop (the largest backward-branch
range) and that loop's opcode histogram.  The dynamic instruction count of one iteration follows the static size of
the (mostly straight-line) loop, so this is the quick check before a variant is timed on the GPU."""
import collections, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = os.path.join(ROOT, "modelpredictivecontrol.jl_b200", "csrc", "warp_inst_11.cu")
KEY = "step_warpILi11ELi2E"


def main(extra):
    with tempfile.TemporaryDirectory() as td:
        obj = os.path.join(td, "w.o")
        r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                            "--expt-extended-lambda", "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + extra + ["-c", "-o", obj, SRC],
                           capture_output=True, text=True)
        if r.returncode:
            sys.exit(r.stderr)
        lines = r.stderr.splitlines()
        for i, l in enumerate(lines):
            if KEY in l and "Compiling" in l:
                print(" | ".join(x.strip() for x in lines[i + 1:i + 4] if "registers" in x or "spill" in x))
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    ins, on = [], False
    for l in sass.splitlines():
        if "Function :" in l:
            on = KEY in l
            continue
        if on:
            m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
            if m:
                ins.append((int(m.group(1), 16), m.group(2)))
    print("kernel instructions:", len(ins))
    addr_ix = {a: i for i, (a, _) in enumerate(ins)}
    best = (0, 0, 0)
    for i, (a, s) in enumerate(ins):
        m = re.search(r"BRA.*?0x([0-9a-f]+)", s)
        if m:
            t = int(m.group(1), 16)
            if t in addr_ix and t < a and i - addr_ix[t] > best[0]:
                best = (i - addr_ix[t], addr_ix[t], i)
    # the outermost backward branch is the instance loop; the interior-point loop is the largest one nested inside it
    spans = []
    for i, (a, s) in enumerate(ins):
        m = re.search(r"BRA.*?0x([0-9a-f]+)", s)
        if m:
            t = int(m.group(1), 16)
            if t in addr_ix and t < a:
                spans.append((i - addr_ix[t] + 1, addr_ix[t], i))
    spans.sort(reverse=True)
    for n, lo, hi in spans[:4]:
        print("backward-branch span: %5d instructions  [%d, %d]" % (n, lo, hi))
    for n, lo, hi in spans:  # the interior-point loop: the largest span without a global store
        ops = [re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in ins[lo:hi + 1]]
        if "STG" in ops:
            continue
        hist = collections.Counter(ops)
        print("interior-point loop: %d instructions [%d, %d]" % (n, lo, hi))
        print("  opcode histogram:", ", ".join("%s %d" % kv for kv in hist.most_common(26)))
        break


if __name__ == "__main__":
    main(sys.argv[1:])
