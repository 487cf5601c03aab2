"""Life 16384^2 after the lean ghost block (the kernel went from issue-bound to latency-bound: ALU pipe 75 % -> 67 %): rows in flight per CTA
(Tuning.prefetch_rows), chunk height and resident CTAs, same box.  `--prebuild` compiles here."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paraiso_b200.build import build_machine  # noqa: E402
from paraiso_b200.examples.life import life_om, life_setup  # noqa: E402

CASES = {
    "a": [(3, 20, 9), (4, 20, 9), (5, 20, 9), (6, 20, 9), (2, 20, 9), (4, 24, 9), (4, 32, 9), (5, 32, 9), (3, 16, 9), (3, 24, 9), (3, 32, 9), (4, 16, 9),
          (4, 20, 8), (5, 24, 8), (4, 20, 10), (3, 20, 10)],
    "b": [(4, 16, 9), (5, 16, 9), (6, 16, 9), (7, 16, 9), (4, 12, 9), (5, 12, 9), (6, 12, 9), (7, 12, 9), (4, 14, 9), (5, 14, 9), (6, 14, 9),
          (4, 10, 9), (5, 10, 9), (6, 10, 9), (5, 8, 9), (6, 8, 9), (6, 18, 9), (7, 20, 9)],
    # (rows in flight, chunk rows, resident CTAs, Tuning.clean_ctas)
    "c": [(6, 16, 9, False), (6, 16, 9, True), (5, 16, 9, True), (7, 16, 9, True), (6, 14, 9, True), (6, 18, 9, True), (6, 20, 9, True), (4, 16, 9, True),
          (6, 16, 8, True), (6, 16, 10, True)],
    # (..., store hint, staging)
    "d": [(6, 16, 9, False, "", "cp_async"), (6, 16, 9, False, "cs", "cp_async"), (6, 16, 9, False, "cg", "cp_async"), (6, 16, 9, False, "", "bulk"),
          (3, 16, 9, False, "", "bulk"), (6, 16, 9, False, "cs", "bulk")],
    # (..., Tuning.warp_rings)
    "e": [(6, 16, 9, False, "", "cp_async", False), (6, 16, 9, False, "", "cp_async", True), (4, 16, 9, False, "", "cp_async", True),
          (8, 16, 9, False, "", "cp_async", True), (6, 20, 9, False, "", "cp_async", True), (6, 12, 9, False, "", "cp_async", True),
          (6, 24, 9, False, "", "cp_async", True), (6, 32, 9, False, "", "cp_async", True), (3, 20, 9, False, "", "cp_async", True),
          (6, 16, 8, False, "", "cp_async", True), (6, 16, 10, False, "", "cp_async", True)],
    "f": [(6, 16, 9), (6, 15, 9), (6, 17, 9), (7, 15, 9), (7, 17, 9), (5, 17, 9), (5, 15, 9), (7, 18, 9), (8, 17, 9)],
}[next((a for a in sys.argv[1:] if a in ("a", "b", "c", "d", "e", "f")), "a")]
size = (16384, 16384)
if __name__ == "__main__":
    built = []
    for case in CASES:
        pf, cr, minb = case[:3]
        clean = case[3] if len(case) > 3 else False
        hint, staging = (case[4], case[5]) if len(case) > 5 else ("", "cp_async")
        warp = case[6] if len(case) > 6 else False
        s = life_setup("master")
        s.tuning.prefetch_rows, s.tuning.chunk_rows_light, s.tuning.min_blocks, s.tuning.clean_ctas = pf, cr, minb, clean
        s.tuning.store_hint, s.tuning.staging, s.tuning.warp_rings = hint, staging, warp
        try:
            built.append(((pf, cr, minb, clean, hint, staging, warp), build_machine(s, life_om("master"), tag=f"variant_Life_pf{pf}_c{cr}_b{minb}_k{int(clean)}_{hint}_{staging}_w{int(warp)}", verbose=True)))
        except Exception as e:
            print(json.dumps(dict(prefetch_rows=pf, chunk_rows=cr, min_blocks=minb, error=repr(e)[:200])), flush=True)
    if "--prebuild" in sys.argv:
        sys.exit(0)
    import torch
    from paraiso_b200.machines import life_seed
    from paraiso_b200.runtime import Machine
    from paraiso_b200.tuning import measure
    seed = torch.from_numpy(life_seed(size[0], 0, size[1])).pin_memory()
    ref = None
    for rep in range(2):
        for (pf, cr, minb, clean, hint, staging, warp), (desc, so) in built:
            try:
                m = Machine(desc, so, size=size)
                m.call("init")
                m.set_from_host("cell", seed)
                st = m.kernels["proceed"]["stages"][0]
                ms = min(measure(m, "proceed", steps=20, stage=0) for _ in range(3))
                m.set_from_host("cell", seed)
                for _ in range(3):
                    m.call("proceed")
                got = (int(m.get("cell").astype("int64").sum()), int(m.scalar("population")))
                ref = ref or got
                print(json.dumps(dict(prefetch_rows=pf, chunk_rows=cr, min_blocks=minb, clean_ctas=clean, store_hint=hint, staging=staging, warp_rings=warp, occupancy=getattr(m.lib, st["symbol"] + "_occupancy")(),
                                      chunks=m._geom(st).nchunks, ms=ms, GBs=2 * 4 * size[0] * size[1] / ms / 1e6, same_result=(got == ref))), flush=True)
                del m
            except Exception as e:
                print(json.dumps(dict(prefetch_rows=pf, chunk_rows=cr, min_blocks=minb, error=repr(e)[:200])), flush=True)
