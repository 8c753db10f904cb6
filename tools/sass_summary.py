"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG/UTMASTG/UBLKCP ...).

    python tools/sass_summary.py [tag]  ->  profiles/<tag>_sass_summary.md   (default tag r02)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "speech_b200", "libspeech_b200.so")
PATTERNS = [("UTCHMMA", r"\bUTCHMMA"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"),
            ("UTMALDG", r"\bUTMALDG"), ("UTMALDG.MULTICAST", r"\bUTMALDG\S*MULTICAST"),
            ("UTMASTG", r"\bUTMASTG"), ("UTMAREDG", r"\bUTMAREDG"), ("UBLKCP", r"\bUBLKCP"),
            ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("HMMA (legacy)", r"\bHMMA"),
            ("MUFU", r"\bMUFU"), ("RED/ATOM", r"\b(RED|ATOM)[GS]?\b")]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", name)
            kernels[cur] = collections.Counter()
            kernels[cur]["_insts"] = 0
            continue
        if cur is None or "/*" not in line:
            continue
        kernels[cur]["_insts"] += 1
        for label, pat in PATTERNS:
            if re.search(pat, line):
                kernels[cur][label] += 1
    cols = [p[0] for p in PATTERNS]
    lines = ["# SASS summary of `speech_b200/libspeech_b200.so` (%s)" % tag, "",
             "`cuobjdump -sass` of the shipped library (sm_100a), instruction counts per kernel of the",
             "mnemonics `/opt/skills/guides/B200_PROFILING.md` lists as evidence: `UTCHMMA` = tcgen05.mma",
             "(`.2CTA` = cta_group::2), `LDTM` = tcgen05.ld, `UTMALDG`/`UTMASTG`/`UTMAREDG` = TMA tensor",
             "load / store / reduce-add, `UBLKCP` = cp.async.bulk (1-D bulk copy), `UTCBAR` = tcgen05.commit,",
             "`SYNCS` = mbarrier ops.  `HMMA` (legacy mma.sync) must be absent.  Regenerate with",
             "`python tools/sass_summary.py %s`." % tag, "",
             "| kernel | SASS instr | " + " | ".join(cols) + " |",
             "|---|---:|" + "---:|" * len(cols)]
    for name, c in kernels.items():
        lines.append("| `%s` | %d | %s |" % (name[:70], c["_insts"],
                                            " | ".join(str(c[x]) if c[x] else "" for x in cols)))
    path = os.path.join(ROOT, "profiles", "%s_sass_summary.md" % tag)
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print(path)


if __name__ == "__main__":
    main()
