"""Static evidence for profiles/: `ptxas -v` resource usage and a SASS opcode histogram per kernel of libemf_b200.so.
  python scripts/sass_report.py r2   ->  profiles/r2_ptxas_sass.md   (no GPU needed: nvcc cross-compiles, cuobjdump reads the .so)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from emfusion_b200 import build as B
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
out = [f"# ptxas -v and SASS opcode histograms ({tag})", "",
       "`nvcc " + " ".join(B.NVCC_FLAGS) + " -Xptxas=-v` per source; `cuobjdump -sass emfusion_b200/lib/libemf_b200.so`.", ""]
# ---- ptxas -v
res = {}
for src in B.SOURCES:
    r = subprocess.run(["nvcc", *B.NVCC_FLAGS, "-Xptxas=-v", "-c", os.path.join(B.CSRC, src), "-o", "/dev/null"], capture_output=True, text=True)
    cur = None
    for ln in r.stderr.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", ln)
        if m:
            cur = m.group(1); res[cur] = {"src": src}
        elif cur and "bytes stack frame" in ln:
            res[cur]["stack"] = ln.strip()
        elif cur and "Used" in ln:
            res[cur]["used"] = ln.split("Used", 1)[1].strip(); cur = None
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
out += ["## Resource usage (`ptxas -v`)", "", "| kernel | source | registers / barriers / smem | stack, spills |", "|---|---|---|---|"]
for k, v in sorted(res.items(), key=lambda kv: kv[1]["src"]):
    out.append(f"| `{demangle(k)}` | {v['src']} | {v.get('used', '')} | {v.get('stack', '')} |")
# ---- SASS
sass = subprocess.run(["cuobjdump", "-sass", B.LIB], capture_output=True, text=True).stdout
hist, name = {}, None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = m.group(1); hist[name] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and name:
        hist[name][m.group(1)] += 1
out += ["", "## SASS (instructions in the binary, not executed counts)", "",
        "| kernel | instructions | LDG.E (128-bit) | STG.E (128-bit) | FFMA+FMUL+FADD | MUFU | SHFL/VOTE/REDUX | ATOM/RED | top opcodes |", "|---|---|---|---|---|---|---|---|---|"]
for k, c in sorted(hist.items(), key=lambda kv: -sum(kv[1].values())):
    d = demangle(k)
    if not d.startswith("emfb::") and "k_" not in d:
        continue
    n = sum(c.values())
    g = lambda p: sum(v for o, v in c.items() if o.startswith(p))
    g128 = lambda p: sum(v for o, v in c.items() if o.startswith(p) and ".128" in o)
    top = ", ".join(f"{o} {v}" for o, v in c.most_common(8))
    out.append(f"| `{d}` | {n} | {g('LDG')} ({g128('LDG')}) | {g('STG')} ({g128('STG')}) | {g('FFMA') + g('FMUL') + g('FADD')} | {g('MUFU')} | "
               f"{g('SHFL') + g('VOTE') + g('REDUX')} | {g('ATOM') + g('RED')} | {top} |")
n_tma = sum(1 for ln in sass.splitlines() if "UTMALDG" in ln or "UTMASTG" in ln)
out += ["", f"TMA instructions (UTMALDG / UTMASTG) in the library: {n_tma} -- the path has no dense tile to stage; DESIGN.md section 9 has the A/B "
        "of shared-memory staging for the integrate's image gathers.", ""]
open(os.path.join(ROOT, "profiles", f"{tag}_ptxas_sass.md"), "w").write("\n".join(out))
print("\n".join(out[:14]))
