"""SASS evidence for profiles/: per kernel resource usage (ptxas -v) and a histogram of the
mnemonics that prove the Blackwell-native path (UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load,
LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier, MULTIMEM / ATOM / RED = collectives),
plus the MMA-issue loop of the hot tensor-core kernels verbatim."""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "hefl_b200", "_obj")
OUT = os.path.join(ROOT, "profiles", "sass")
KEY = re.compile(r"\b(LDSM|UTCHMMA|UTCQMMA|UTCCP|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|UTCBAR|UTCATOMSWS|SYNCS|LDGMC|STGMC|REDGMC|MULTIMEM|ATOMG|ATOM|REDG|RED|LDGSTS|"
                 r"HMMA|IMAD|LDS|STS|LDG|STG|SHFL|BAR|MEMBAR|ERRBAR|CCTL|ELECT|R2UR|UIADD3|UMOV|FFMA|FFMA2|FHFMA|UCGABAR_ARV|UCGABAR_WAIT|"
                 r"ACQBULK|DFMA|DADD|DMUL)\b")


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    return r.stdout.splitlines()


def main():
    os.makedirs(OUT, exist_ok=True)
    for obj in sorted(glob.glob(os.path.join(OBJ, "*.cu.o"))):
        base = os.path.basename(obj).replace(".cu.o", "")
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        log = open(obj + ".log").read() if os.path.exists(obj + ".log") else ""
        res = {}
        cur = None
        for line in log.splitlines():
            m = re.search(r"Compiling entry function '([^']+)'", line)
            if m:
                cur = m.group(1)
            m = re.search(r"Used (\d+) registers.*?(?:, (\d+) bytes smem)?", line)
            if m and cur:
                res[cur] = line.split("ptxas info    : ")[-1].strip()
        kernels = collections.OrderedDict()
        name = None
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                name = m.group(1)
                kernels[name] = []
                continue
            if name and re.search(r"/\*[0-9a-f]{4}\*/", line):
                kernels[name].append(line)
        names = list(kernels)
        pretty = dict(zip(names, demangle(names)))
        md = [f"# SASS summary — `csrc/{base.replace('_', '/', 1)}.cu` (sm_100a, `cuobjdump -sass`)\n"]
        for k, lines in kernels.items():
            hist = collections.Counter()
            for l in lines:
                ins = re.sub(r"/\*[0-9a-f]+\*/", "", l)
                for m in KEY.finditer(ins):
                    hist[m.group(1)] += 1
                    break
            md.append(f"\n## `{pretty[k][:150]}`\n")
            md.append(f"- instructions: {len(lines)}; ptxas: {res.get(k, 'n/a')}")
            md.append("- mnemonics: " + ", ".join(f"{a}×{b}" for a, b in hist.most_common(16)))
            verb = [re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() for l in lines
                    if re.search(r"LDGMC|STGMC|UTCQMMA|UTCCP|UTMALDG\.2D\.MULTICAST|UTMASTG|UBLKCP|UCGABAR_ARV|LDG\.E\.ENL2\.256", l)]
            if verb:
                md.append("\nVerbatim (first of each kind):\n\n```")
                shown = set()
                for l in verb:
                    k = re.search(r"(LDGMC|STGMC|UTCQMMA|UTCCP|UTMALDG\.2D\.MULTICAST|UTMASTG|UBLKCP|UCGABAR_ARV|LDG\.E\.ENL2\.256)", l).group(1)
                    if k not in shown:
                        shown.add(k)
                        md.append(l)
                md.append("```")
            if "UTCHMMA" in hist and ("tap_gemm_kernel<16, 32, true>" in pretty[k].replace("(int)", "").replace("(bool)1", "true")
                                      or "wgrad_kernel<16, 32, false>" in pretty[k].replace("(int)", "").replace("(bool)0", "false")):
                idx = [i for i, l in enumerate(lines) if "UTCHMMA" in l]
                lo, hi = max(0, idx[0] - 12), min(len(lines), idx[min(len(idx) - 1, 7)] + 6)
                md.append("\nMMA issue loop (first tcgen05.mma instructions, verbatim):\n\n```")
                md += [re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() for l in lines[lo:hi]]
                md.append("```")
        open(os.path.join(OUT, base + ".md"), "w").write("\n".join(md) + "\n")
        print("wrote", os.path.join("profiles/sass", base + ".md"), len(kernels), "kernels")


if __name__ == "__main__":
    main()
