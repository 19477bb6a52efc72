#!/usr/bin/env python
"""Join an ncu report's SASS page with nvdisasm line info and aggregate executed instructions and
stall samples per CUDA source line (ncu's CLI cannot export the per-line view as CSV).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep motcpp_b200/libmotb200.so bytetrack_step_kernel [top]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def line_map(so, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    txt = ""
    for f in sorted(os.listdir(tmp)):          # the library is several translation units: find the cubin that holds the kernel
        if not f.endswith(".cubin"):
            continue
        t = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if re.search(r"\.text\.[^\n]*" + re.escape(kernel), t):
            txt = t
            break
    m = {}
    chain = []
    fresh = True
    inside = False
    for ln in txt.splitlines():
        if ln.startswith("//---------------------"):
            inside = (".text." in ln) and (kernel in ln)
            continue
        if not inside:
            continue
        f = re.match(r'\s*//## File "(.*?)", line (\d+)', ln)
        if f:
            if fresh:
                chain = []
                fresh = False
            chain.append((os.path.basename(f.group(1)), int(f.group(2))))
            continue
        a = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(\S.*?);", ln)
        if a and chain:
            m[int(a.group(1), 16)] = tuple(chain)      # innermost first ... outermost (kernel body) last
            fresh = True
    return m


def user_leaf(chain):
    for fr in chain:
        if not fr[0].endswith(".hpp") and not fr[0].endswith(".h"):
            return fr
    return chain[0]


def phase(chain, phase_file):
    """the frame of `phase_file` closest to the kernel body (i.e. the call site inside the frame function)"""
    cands = [fr for fr in chain if fr[0] == phase_file]
    if len(cands) >= 2:
        return cands[-2]
    return cands[-1] if cands else chain[-1]


PHASE_FILE = os.environ.get("NCU_PHASE_FILE", "bytetrack_kernel.cuh")


def main():
    rep, so, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    lm = line_map(so, kernel)        # `kernel` is matched against the MANGLED name in the cubin
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in rows if r and r[0].startswith("0x")]
    base = min(int(r[ia], 16) for r in data)
    agg = {}
    phases = {}
    tot_i = tot_s = 0
    for r in data:
        off = int(r[ia], 16) - base
        ch = lm.get(off, (("?", 0),))
        key = user_leaf(ch)
        ph = phase(ch, PHASE_FILE)
        pa = phases.setdefault(ph, [0, 0])
        inst, samp = int(r[ii] or 0), int(r[isamp] or 0)
        a = agg.setdefault(key, [0, 0, {}])
        a[0] += inst
        a[1] += samp
        pa[0] += inst
        pa[1] += samp
        for c in stall_cols:
            v = int(r[c] or 0)
            if v:
                a[2][hdr[c]] = a[2].get(hdr[c], 0) + v
        tot_i += inst
        tot_s += samp
    print(f"total warp instructions {tot_i:,}  samples {tot_s:,}")
    print("== by instructions executed")
    for key, (inst, samp, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{key[0]}:{key[1]:<5} inst {100*inst/tot_i:5.1f}%  samples {100*samp/max(tot_s,1):5.1f}%")
    print("== by stall samples")
    for key, (inst, samp, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        tops = ", ".join(f"{k[6:]} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{key[0]}:{key[1]:<5} samples {100*samp/max(tot_s,1):5.1f}%  inst {100*inst/tot_i:5.1f}%  [{tops}]")
    print("== by call site in " + PHASE_FILE + " (inlined callees included)")
    for key, (inst, samp) in sorted(phases.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{key[0]}:{key[1]:<5} inst {100*inst/tot_i:5.1f}%  samples {100*samp/max(tot_s,1):5.1f}%")
    # per-file totals
    files = {}
    for key, (inst, samp, st) in agg.items():
        f = files.setdefault(key[0], [0, 0])
        f[0] += inst
        f[1] += samp
    print("== by file")
    for f, (inst, samp) in sorted(files.items(), key=lambda kv: -kv[1][0]):
        print(f"{f:<28} inst {100*inst/tot_i:5.1f}%  samples {100*samp/max(tot_s,1):5.1f}%")


if __name__ == "__main__":
    main()
