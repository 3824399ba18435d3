"""One block of headline numbers per kernel of an .ncu-rep (multi-kernel reports): duration, DRAM bytes, issue / pipe
utilisation, occupancy, registers, warp instructions, top stall reasons.  usage: python tools/ncu_kernels.py <rep>"""
import csv
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], stdout=subprocess.PIPE, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    names, units, rows = raw[0], raw[1], raw[2:]
    ix = {n: i for i, n in enumerate(names)}
    want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
            ("smsp__inst_executed.sum", "warp instructions"), ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
            ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
            ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %"),
            ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu pipe %"),
            ("sm__instruction_throughput.avg.pct_of_peak_sustained_active", "issue %"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
            ("launch__registers_per_thread", "registers/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
            ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
            ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
            ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %")]
    src = page(rep, "source")
    starts = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"] + [len(src)]
    stall_by_kernel = {}
    for a, b in zip(starts[:-1], starts[1:]):
        h = src[a + 1]
        hx = {n: i for i, n in enumerate(h)}
        stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        tot = {}
        for r in src[a + 2:b]:
            for s in stalls:
                try:
                    tot[s] = tot.get(s, 0.0) + float(r[hx[s]])
                except Exception:
                    pass
        t = sum(tot.values()) or 1.0
        stall_by_kernel[src[a][1]] = ", ".join("%s %.1f%%" % (k[6:], 100 * v / t) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:7])
    for r in rows:
        kname = r[ix["Kernel Name"]] if "Kernel Name" in ix else "?"
        print("== %s" % kname)
        for key, label in want:
            if key in ix:
                print("   %-24s %s %s" % (label, r[ix[key]], units[ix[key]]))
        base = kname.replace("void ", "").replace("trm::", "").split("(")[0]
        for k, v in stall_by_kernel.items():
            if k.replace("void ", "").replace("trm::", "").split("(")[0] == base:
                print("   stalls: " + v)


if __name__ == "__main__":
    main()
