"""Host logic of the bucketed data-parallel step (no GPU): Engine._dp_plan cuts the towers' gradient arena into
buckets in the order the backward pass completes them; every float and every frozen-BN row belongs to exactly one."""
import pytest

from monopsr_b200.core import model_spec as ms
from monopsr_b200.core.engine import Engine


def layout_only_engine():
    e = Engine.__new__(Engine)
    e._build_param_layout()
    e.towers, e.bn_row0, row = {}, {}, 0
    for enc in ms.ENCODERS:
        units, cin = [], 64
        for name, base, nunits, rate in ms.BLOCKS:
            for u in range(1, nunits + 1):
                units.append(dict(scope="%s/resnet_v1_101/%s/unit_%d/bottleneck_v1" % (enc, name, u), proj=(cin != base * 4)))
                cin = base * 4
        e.towers[enc] = dict(units=units)
        for scope, k, cin_, cout, _ in ms.conv_layers(enc):
            e.bn_row0[scope] = row
            row += cout
    e.bn_rows, e.bn_layers = row, object()
    e.round_off = min(e.layout[n][1] for n in e.trainable_names if not n.startswith("FirstStage"))
    return e


@pytest.mark.parametrize("nb", [1, 2, 3, 4, 6, 8])
def test_buckets_partition_the_tower_gradients(nb):
    e = layout_only_engine()
    plan = e._dp_plan(nb)
    assert 1 <= len(plan) <= nb and plan[-1]["unit"] == -1
    units = [b["unit"] for b in plan[:-1]]
    assert units == sorted(units, reverse=True)            # completion order of the backward pass: last units first
    for key, end in (("ranges", e.round_off), ("rows", e.bn_rows)):
        cover = sorted(r for b in plan for r in b[key])
        assert cover[0][0] == 0 and cover[-1][1] == end
        assert all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
    if nb == 4:
        sizes = [sum(r[1] - r[0] for r in b["ranges"]) for b in plan]
        assert max(sizes) < 1.3 * (sum(sizes) / 4)          # balanced: the exposed last bucket is small
        assert sizes[-1] <= min(sizes[:-1])


def test_bucket_boundaries_are_unit_boundaries():
    e = layout_only_engine()
    for b in e._dp_plan(4)[:-1]:
        U = e.towers[ms.ENCODERS[0]]["units"][b["unit"]]
        first = U["scope"] + ("/shortcut" if U["proj"] else "/conv1") + "/weights"
        assert b["ranges"][0][0] == e.layout[first][1]


def test_tower_variables_are_two_contiguous_runs_of_the_trainable_list():
    """Engine.opt_parts['tower0' / 'tower1'] (data parallelism: one all-reduce and one train-op per tower) relies on the
    trainable variables being ordered [first tower | second tower | everything else]"""
    from monopsr_b200.core import model_spec as ms
    names = [n for n, s, k in ms.param_table() if k in ms.TRAINABLE_KINDS]
    t_head = min(i for i, n in enumerate(names) if not n.startswith("FirstStage"))
    assert all(not n.startswith("FirstStage") for n in names[t_head:])
    first = names[0].split("/")[0]
    t_mid = min(i for i, n in enumerate(names[:t_head]) if n.split("/")[0] != first)
    assert all(n.split("/")[0] == first for n in names[:t_mid])
    assert len({n.split("/")[0] for n in names[t_mid:t_head]}) == 1
    assert {first, names[t_mid].split("/")[0]} == set(ms.ENCODERS)
    assert 0 < t_mid < t_head < len(names)
