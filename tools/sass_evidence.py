"""Which SASS instructions each kernel of libprime_match.so uses for its data movement (no GPU needed):
UBLKCP = cp.async.bulk (1-D TMA), SYNCS.* = mbarrier, LDG/STG .128 = 128-bit vector accesses
(.EF = evict-first streaming stores), ATOMG/RED = global atomics.  Template instances are folded per kernel
name (max over instances)."""
import collections, os, re, subprocess, sys

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "protocol_b200", "libprime_match.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pat = re.compile(r"\b(UBLKCP[.\w]*|SYNCS[.\w]*|LDG\.E(?:\.\w+)*?\.128[.\w]*|LDG\.E(?:\.\w+)*?\.64[.\w]*|STG\.E(?:\.EF)?\.128|STG\.E(?:\.EF)?\.64|ATOMG[.\w]*|RED[.\w]*|ATOMS[.\w]*|SHFL[.\w]*|VOTE[.\w]*|LDS\.128)")
per = collections.defaultdict(lambda: collections.Counter())
fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        fn = re.sub(r"<.*", "", name).replace("void ", "").replace("pm::", "")
        inst = collections.Counter()
        per[fn]["__instances__"] += 1
        cur = inst
        continue
    if fn is None or not fn.startswith("pm_"):
        continue
    m = pat.search(line)
    if m:
        key = m.group(1).split(".")[0] if m.group(1).startswith(("SHFL", "VOTE", "ATOMS")) else m.group(1)
        per[fn][key] += 1
rows = []
for fn, c in sorted(per.items()):
    if not fn.startswith("pm_"):
        continue
    n = c.pop("__instances__")
    rows.append((fn, n, ", ".join(f"{k} x{v // n if v >= n else v}" for k, v in sorted(c.items()))))
w = max(len(r[0]) for r in rows)
print(f"{'kernel'.ljust(w)}  inst  per-instance SASS (count)")
for fn, n, s in rows:
    print(f"{fn.ljust(w)}  {n:>4}  {s}")
