"""profiles/r2_sass_{cbfqp,odcbf,mpc}.txt: cuobjdump -sass of the shipped libscb.so, per hot kernel: instruction mix,
Blackwell / Hopper-class async instructions (UBLKCP = cp.async.bulk, SYNCS = mbarrier), and an excerpt of the hot loop.
    python tools/sass_excerpts.py            (runs on the GPU-less build box)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "safe_control_b200", "libscb.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
funcs = {}
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        funcs[cur].append(line)
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
WANT = {"cbfqp": ["cbfqp_tma_kernelILi1ELi4ELi5E", "cbfqp_kernelILi1ELi32ELi1ELb1E"], "odcbf": ["odcbf_tma_kernelILi3ELi1ELi8ELi4E"],
        "mpc": ["mpc_kernelILi1ELi32E", "mpc_kernelILi4ELi32E"]}
op = lambda l: re.sub(r"^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+\s+)?", "", l).split()[0].rstrip(";")
for tag, pats in WANT.items():
    out = [f"# SASS of the {tag} hot kernels in safe_control_b200/libscb.so (cuobjdump -sass, sm_100a); tools/sass_excerpts.py", ""]
    for pat in pats:
        for name, lines in funcs.items():
            if pat not in name:
                continue
            mix = collections.Counter(op(l).split(".")[0] for l in lines)
            out.append(f"## {demangle(name)[:160]}")
            out.append(f"instructions: {len(lines)}")
            out.append("mix: " + ", ".join(f"{k} {v}" for k, v in mix.most_common(24)))
            asyncs = [l.strip() for l in lines if re.search(r"UBLKCP|SYNCS|UTMA|FENCE\.VIEW\.ASYNC|LDGSTS|ELECT", l)]
            out.append(f"async-copy / mbarrier instructions ({len(asyncs)}):")
            out += ["    " + re.sub(r"\s+", " ", a)[:150] for a in asyncs[:16]]
            wide = sum(1 for l in lines if re.search(r"LDG\.E\.128|LDS\.128|LDG\.E\.64|LDS\.64", l))
            out.append(f"64/128-bit loads: {wide};  LDL {mix.get('LDL', 0)}  STL {mix.get('STL', 0)};  DFMA {mix.get('DFMA', 0)}  DMUL {mix.get('DMUL', 0)}  DADD {mix.get('DADD', 0)}")
            # excerpt: 60 lines around the first bulk copy (QP kernels) or the first DFMA-dense stretch (MPC)
            idx = next((i for i, l in enumerate(lines) if "UBLKCP" in l), None)
            if idx is None:
                dens = [sum(1 for l in lines[i:i + 40] if "DFMA" in l) for i in range(0, max(1, len(lines) - 40))]
                idx = max(range(len(dens)), key=dens.__getitem__) if dens else 0
            lo = max(0, idx - 20)
            out.append(f"excerpt (instructions {lo}..{lo + 60}):")
            out += ["    " + re.sub(r"\s+", " ", l.strip())[:160] for l in lines[lo:lo + 60]]
            out.append("")
    open(os.path.join(ROOT, "profiles", f"r2_sass_{tag}.txt"), "w").write("\n".join(out) + "\n")
    print(tag, "written", len(out), "lines")
