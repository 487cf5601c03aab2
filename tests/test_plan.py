"""The re-implemented analysis passes reproduce the reference's plan for the checked-in samples:
storage (manifest) node ids, subkernel count and margins as visible in
examples-old/Life-exampled/dist/Life.hpp:20-26,58-61 and examples-old/Hydro-exampled/dist/Hydro.hpp:27-77,111-120."""
from paraiso_b200 import annotation as A
from paraiso_b200.examples.hydro import hydro_om, hydro_setup
from paraiso_b200.examples.life import life_om, life_setup
from paraiso_b200.generator.plan import stencil_radius, translate
from paraiso_b200.optimization import optimize

HYDRO_MANIFESTS = [(0, 8), (0, 10), (0, 12), (0, 46), (0, 52), (0, 86), (0, 93), (1, 17), (1, 19), (1, 72), (1, 74), (1, 76),
                   (1, 308), (1, 310), (1, 312), (1, 322), (1, 324), (1, 326), (1, 336), (1, 338), (1, 340), (1, 342), (1, 344),
                   (1, 348), (1, 352), (1, 358), (1, 362), (1, 366), (1, 370), (1, 376), (1, 1064), (1, 1068), (1, 1072),
                   (1, 1078), (1, 1082), (1, 1086), (1, 1090), (1, 1096), (1, 1798), (1, 1806), (1, 1814), (1, 1822),
                   (1, 3846), (1, 3854), (1, 3862), (1, 3870), (1, 3912), (1, 3914), (1, 3946), (1, 3950), (1, 3952)]


def test_life_exampled_plan():
    p = translate(life_setup("exampled"), life_om("exampled"))
    assert [s.manifest for s in p.storages if s.manifest] == [(0, 99), (0, 102), (0, 105), (1, 67), (1, 69), (1, 74)]
    assert [(s.name, s.realm) for s in p.sub_kernels] == [("om_init_sub_0", "Array"), ("om_init_sub_1", "Scalar"),
                                                          ("om_proceed_sub_2", "Array"), ("om_proceed_sub_3", "Scalar")]
    assert p.lower_margin == (1, 1) and p.upper_margin == (1, 1) and p.memory_size == (130, 130)


def test_hydro_exampled_plan():
    p = translate(hydro_setup((1024, 1024)), hydro_om("exampled"))
    assert [s.manifest for s in p.storages if s.manifest] == HYDRO_MANIFESTS
    assert len(p.sub_kernels) == 10
    assert [len(k.dataflow.nodes) for k in p.om.kernels] == [95, 3958]   # ~3,958 nodes
    assert p.lower_margin == (3, 3) and p.upper_margin == (3, 3) and p.memory_size == (1030, 1030)
    main = p.sub_kernels[8]
    assert (len(main.input_idxs), len(main.output_idxs)) == (27, 4)      # Hydro_sub_8 27 -> 4


def test_life_master_plan_matches_appendix_a():
    """Cyclic 80x48 -> margins 0, init = sub_0 (Array) + sub_1 (Scalar), proceed = sub_2 + sub_3."""
    p = translate(life_setup("master"), life_om("master"))
    assert p.lower_margin == (0, 0) and p.upper_margin == (0, 0) and p.memory_size == (80, 48)
    assert [(s.name, s.realm) for s in p.sub_kernels] == [("om_init_sub_0", "Array"), ("om_init_sub_1", "Scalar"),
                                                          ("om_proceed_sub_2", "Array"), ("om_proceed_sub_3", "Scalar")]
    assert [s.name for s in p.storages[:3]] == ["om_s0_cell", "om_s1_population", "om_s2_generation"]
    assert stencil_radius(life_om("master")) == ((1, 1), (1, 1))


def test_master_hydro_subkernel_order_respects_dependencies():
    """Every input of a subkernel that is a Manifest node is produced by an earlier subkernel
    (the reference's greedy grouping violates this for `broadcast $ cast $ loadSize`; see optimization.py)."""
    p = translate(hydro_setup((64, 64)), hydro_om("master"))
    produced = {}
    for s in p.sub_kernels:
        for i in s.input_idxs:
            nd = p.om.kernels[s.kernel_idx].dataflow.nodes[i]
            if A.to_maybe(A.Allocation, nd.anot) == A.Manifest:
                assert produced[(s.kernel_idx, i)] < s.om_write_group_idx
        for o in s.output_idxs:
            produced[(s.kernel_idx, o)] = s.om_write_group_idx


def test_optimize_is_idempotent_by_level():
    om = life_om("master")
    optimize("O3", om)
    n = [len(k.dataflow.nodes) for k in om.kernels]
    optimize("O3", om)
    assert n == [len(k.dataflow.nodes) for k in om.kernels]
