#!/usr/bin/env python
"""Per-stage table of the CMDG_TIMELINE dumps (one CSV per rank; libcmdg records labelled CUDA events on
the stream that does the work during the last two steps of a timed cmdg_lsrk_steps call).

    CMDG_TIMELINE=gpurun_out/tl python -m torch.distributed.run ... bench.py --gpus N --headline-only
    python tools/timeline_table.py gpurun_out/tl > profiles/r2_timeline_nN.md

Columns (microseconds, last recorded step): for every stage the exterior chain on the side stream
(exterior kernel, pack, wait for NCCL to start, NCCL send/recv, unpack) next to the interior kernel on
the main stream.  The two chains are only coupled through kernel-to-kernel events (interior(s+1) needs
exterior-kernel(s), exterior-kernel(s+1) needs interior(s) and its own unpack(s)), so on a fast rank the
interior chain runs up to one stage ahead of the exterior chain, which waits in NCCL for the slowest
neighbour: `int_lead` = how far the interior kernel of the stage started before the exterior kernel,
`int_gap` = idle time on the main stream between consecutive interior kernels (what a stage really costs
beyond its interior kernel), `nccl` includes the wait for the neighbours."""
import csv
import glob
import sys
from collections import defaultdict


def load(path):
    rows = list(csv.DictReader(open(path)))
    steps = sorted({int(r["step"]) for r in rows})
    last = steps[-1]
    by_stage = defaultdict(list)
    for r in rows:
        if int(r["step"]) == last:
            by_stage[int(r["stage"])].append((r["label"], int(r["info"]), float(r["t_us"])))
    return by_stage


def stage_row(marks):
    kb = [(i, t) for (l, i, t) in marks if l == "kernel_begin"]
    ke = [(i, t) for (l, i, t) in marks if l == "kernel_end"]
    get = lambda name: [t for (l, i, t) in marks if l == name]
    if len(kb) < 2:      # serial schedule or single rank
        return None
    # exterior launch = the smaller one
    (ie, te0), (ii, ti0) = sorted(kb)[0], sorted(kb)[-1]
    te1 = [t for (i, t) in ke if i == ie][0]
    ti1 = [t for (i, t) in ke if i == ii][0]
    pack, nb, ne_, ub, ue = get("pack_end")[0], get("nccl_begin")[0], get("nccl_end")[0], get("unpack_begin")[0], get("unpack_end")[0]
    start = min(te0, ti0)
    end = max(ue, ti1)
    return dict(ext_kernel=te1 - te0, pack=pack - te1, nccl_wait=nb - pack, nccl=ne_ - nb, unpack_wait=ub - ne_,
                unpack=ue - ub, ext_chain=ue - te0, interior=ti1 - ti0, int_lead=te0 - ti0,
                start=start, end=end, n_ext=ie, n_int=ii, ti0=ti0, ti1=ti1, te0=te0, ue=ue)


def main():
    prefix = sys.argv[1]
    files = sorted(glob.glob(prefix + ".rank*.csv"))
    print(f"# Stage timeline, {len(files)} rank(s), last recorded step (microseconds)\n")
    cols = ["ext_kernel", "pack", "nccl_wait", "nccl", "unpack_wait", "unpack", "ext_chain", "interior",
            "int_lead", "int_gap"]
    summary = []
    for f in files:
        rank = f.split(".rank")[-1].split(".")[0]
        st = load(f)
        rows = [stage_row(st[s]) for s in sorted(st)]
        if any(r is None for r in rows):
            print(f"rank {rank}: serial schedule / single rank (no exterior chain recorded)\n")
            continue
        for a, b in zip(rows, rows[1:]):
            a["int_gap"] = b["ti0"] - a["ti1"]
        rows[-1]["int_gap"] = float("nan")
        print(f"## rank {rank}  (exterior {rows[0]['n_ext']} / interior {rows[0]['n_int']} elements)\n")
        print("| stage | " + " | ".join(cols) + " |")
        print("|---|" + "---|" * len(cols))
        for s, r in enumerate(rows):
            print(f"| {s + 1} | " + " | ".join(f"{r[c]:.1f}" for c in cols) + " |")
        tot = {c: sum(r[c] for r in rows if r[c] == r[c]) for c in cols}
        print("| sum | " + " | ".join(f"{tot[c]:.1f}" for c in cols) + " |\n")
        summary.append((rank, tot))
    if summary:
        print("## per-rank sums over the 5 stages\n")
        print("| rank | interior kernels | gaps between them | exterior kernels | pack + unpack | nccl (incl. waiting for neighbours) |")
        print("|---|---|---|---|---|---|")
        for rank, t in summary:
            print(f"| {rank} | {t['interior']:.1f} | {t['int_gap']:.1f} | {t['ext_kernel']:.1f} | {t['pack'] + t['unpack']:.1f} | {t['nccl']:.1f} |")


FAMILY = {1: "filter", 2: "gradient", 3: "column", 4: "tendency"}


def ocean_main():
    """`timeline_table.py --ocean <prefix>`: the HBModel marks carry `kernel family * 1e8 + elements` in `info`
    (hb_eval_t).  Prints, per rank and for the middle stage of the last recorded step, every kernel / halo phase with
    begin, end and duration, and the per-family sums over the whole step."""
    prefix = sys.argv[2]
    files = sorted(glob.glob(prefix + ".rank*.csv"))
    print(f"# Ocean stage timeline, {len(files)} rank(s), last recorded step (microseconds)\n")
    for f in files:
        rank = f.split(".rank")[-1].split(".")[0]
        st = load(f)
        stages = sorted(st)
        mid = stages[len(stages) // 2]
        print(f"## rank {rank}, stage {mid + 1} of {len(stages)}\n")
        print("| phase | elements / values | begin | end | us |")
        print("|---|---|---|---|---|")
        marks = st[mid]
        opened = {}
        rows = []
        for lab, info, t in marks:
            if lab == "kernel_begin":
                opened[info] = t
            elif lab == "kernel_end" and info in opened:
                rows.append((opened.pop(info), t, FAMILY.get(info // 100000000, "kernel"), info % 100000000))
        t_of = lambda name: [t for (l, i, t) in marks if l == name]
        for b, e in zip(t_of("nccl_begin"), t_of("nccl_end")):
            rows.append((b, e, "NCCL send/recv", 0))
        for b, e in zip(t_of("unpack_begin"), t_of("unpack_end")):
            rows.append((b, e, "unpack", 0))
        for (l, i, t) in marks:
            if l == "pack_end":
                rows.append((t, t, "pack done", i))
        for b, e, name, n in sorted(rows):
            print(f"| {name} | {n or ''} | {b:.0f} | {e:.0f} | {e - b:.0f} |")
        tot = defaultdict(float)
        for s_ in stages:
            op = {}
            for lab, info, t in st[s_]:
                if lab == "kernel_begin":
                    op[info] = t
                elif lab == "kernel_end" and info in op:
                    tot[FAMILY.get(info // 100000000, "kernel")] += t - op.pop(info)
        t_all = [t for s_ in stages for (_, _, t) in st[s_]]
        print(f"\nstep: {max(t_all) - min(t_all):.0f} us over {len(stages)} stages; kernel sums per family: "
              + ", ".join(f"{k} {v:.0f}" for k, v in sorted(tot.items())) + "\n")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--ocean":
        ocean_main()
    else:
        main()
