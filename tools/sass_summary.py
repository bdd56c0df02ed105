"""Per-kernel Blackwell instruction evidence from the built library (runs anywhere: cuobjdump only).
For every kernel of unispeech_b200/lib/libunispeech_b200.so: counts of UTC*MMA (tcgen05.mma; .2CTA = cta_group::2), LDTM / STTM
(tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA loads / stores / reductions; .MULTICAST), MUFU.EX2, REDG and the legacy HMMA
path (must be 0).  usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, hashlib, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "unispeech_b200", "lib", "libunispeech_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = [("UTCMMA", r"\bUTC[A-Z]*MMA"), ("UTCMMA.2CTA", r"\bUTC[A-Z]*MMA\.2CTA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
        ("UTMALDG", r"\bUTMALDG"), ("UTMALDG.MC", r"\bUTMALDG[.\w]*MULTICAST"), ("UTMASTG", r"\bUTMASTG"), ("UTMAREDG", r"\bUTMAREDG"),
        ("MUFU.EX2", r"\bMUFU\.EX2"), ("REDG", r"\bREDG"), ("HMMA", r"\bHMMA"), ("LDL/STL", r"\b(LDL|STL)\b")]
kern, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("b200::", "").replace("(anonymous namespace)::", "")
        kern[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for name, p in pats:
        if re.search(p, line):
            kern[cur][name] += 1
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}  (sha256 {hashlib.sha256(open(lib,'rb').read()).hexdigest()[:16]}, sm_100a)")
print(f"# {len(kern)} kernels; columns: " + " ".join(n for n, _ in pats))
tot = collections.Counter()
for k, c in kern.items():
    tot.update(c)
    print(f"{k[:96]:96s} " + " ".join(f"{c.get(n, 0):5d}" for n, _ in pats))
print(f"{'TOTAL':96s} " + " ".join(f"{tot.get(n, 0):5d}" for n, _ in pats))
