"""SASS opcode census of libdfine_b200.so: static counts of the Blackwell-native opcodes per kernel.

    python tools/sass_census.py > profiles/r5_sass_opcodes.md
"""
import collections
import re
import subprocess
import sys

LIB = "d-fine-seg_b200/dfine_b200/_C/libdfine_b200.so"
WANT = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "REDG", "ATOMS", "ATOMG", "FFMA2", "HMMA",
        "SHFL", "SYNCS", "DADD", "MUFU.EX2", "MUFU.RCP", "MATCH", "VOTE"]


def strip_params(name: str) -> str:
    """Drop the parameter list of a demangled kernel name (template arguments such as `(int)1` stay)."""
    depth = 0
    for i, ch in enumerate(name):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            return name[:i]
    return name


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.split("\n")
    counts, cur, k = collections.OrderedDict(), None, 0
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = names[k]
            k += 1
            cur = strip_params(cur.replace("void ", ""))
            counts[cur] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P[T0-9]+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            for w in WANT:
                if op.startswith(w):
                    counts[cur][w] += 1
    print("# SASS opcode census of libdfine_b200.so (round 2, final build)\n")
    print("`cuobjdump -sass " + LIB + "`, static instruction counts per kernel (sm_100a; `tools/sass_census.py`).  "
          "`UTCHMMA` = tcgen05.mma, `LDTM` / `STTM` = tcgen05.ld / .st, `UTMALDG` / `UTMASTG` = TMA tensor load / store, "
          "`UBLKCP` = cp.async.bulk, `REDG` = red.global (incl. the multimem.red of `multicast_add_kernel`), `ATOMS` = "
          "shared-memory integer atomics, `FFMA2` = packed fp32 FMA, `DADD` = float64 (dfine_lsap duals).  No `HMMA` "
          "(legacy mma.sync) anywhere.\n")
    print("| kernel | opcodes |\n|---|---|")
    for name, c in counts.items():
        if c:
            print(f"| `{name}` | " + ", ".join(f"{w} x{c[w]}" for w in WANT if c[w]) + " |")
    assert not any(c["HMMA"] for c in counts.values())


if __name__ == "__main__":
    main()
