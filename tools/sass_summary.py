"""profiles/sass_summary.md: per-kernel counts of the SASS mnemonics that prove tcgen05 / TMA / TMEM use (cuobjdump -sass
of the built library; runs without a GPU)."""
import collections
import re
import subprocess
import sys

LIB = "seed-story_b200/lib/libseedstory_b200.so"
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "MUFU.EX2", "LDG", "STG"]


def main(out):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.split("\n")
    blocks = re.split(r"\n\s*Function : \S+\n", txt)[1:]
    rows = []
    for name, body in zip(names, blocks):
        c = collections.Counter()
        for m in re.finditer(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body, flags=re.M):
            op = m.group(1)
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    c[k] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                c["UTCHMMA.2CTA"] += 1
        short = re.sub(r"\(anonymous namespace\)::", "", name)
        short = re.sub(r"\(.*", "", short).replace("void ", "")
        rows.append((short, c, len(re.findall(r"^\s+/\*[0-9a-f]{4,}\*/", body, flags=re.M))))
    rows.sort(key=lambda r: (-(r[1]["UTCHMMA"]), r[0]))
    with open(out, "w") as f:
        f.write("# SASS summary of libseedstory_b200.so (sm_100a)\n\n`python tools/sass_summary.py` = `cuobjdump -sass` of the built library, "
                "instruction mnemonics counted per kernel.\nUTCHMMA = tcgen05.mma (`.2CTA` = cta_group::2), UTMALDG / UTMASTG = "
                "TMA tensor load / store, UBLKCP = cp.async.bulk (1-D), LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, "
                "SYNCS = mbarrier ops, HMMA = mma.sync.\n\n")
        f.write("| kernel | SASS instrs | " + " | ".join(KEYS) + " |\n|---|---|" + "---|" * len(KEYS) + "\n")
        for short, c, n in rows:
            f.write(f"| `{short}` | {n} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |\n")
    print(f"{len(rows)} kernels -> {out}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "profiles/sass_summary.md")
