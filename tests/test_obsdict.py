"""``ObsDict`` (the drop-in ``Environment.get_obs`` result, environment.py:110-130): a real dict whose per-agent
entries are materialised on access -- same contents, order and protocol as the eager dict it replaces."""
import copy
import pickle


def _make(n, log):
    from marl_demandresponse_b200.environment import ObsDict

    def build(i):
        log.append(i)
        return {"id": i, "message": [{"k": i}]}

    return ObsDict(build, n)


def test_entries_are_built_on_access_and_only_once():
    log = []
    o = _make(6, log)
    assert len(o) == 6 and 4 in o and 6 not in o and -1 not in o and "a" not in o and log == []
    assert o[4]["id"] == 4 and o[4] is o[4] and log == [4]
    assert o.get(2)["id"] == 2 and o.get(17) is None and o.get(17, 5) == 5
    try:
        o[9]
        raise AssertionError("out-of-range id must raise KeyError")
    except KeyError:
        pass


def test_whole_dict_views_fill_in_id_order():
    log = []
    o = _make(5, log)
    o[3]                                   # touched out of order first
    assert list(o) == [0, 1, 2, 3, 4] and list(o.keys()) == [0, 1, 2, 3, 4]
    assert [v["id"] for v in o.values()] == [0, 1, 2, 3, 4] and [k for k, _ in o.items()] == [0, 1, 2, 3, 4]
    assert sorted(log) == [0, 1, 2, 3, 4]  # every entry built exactly once
    want = {i: {"id": i, "message": [{"k": i}]} for i in range(5)}
    assert o == want and dict(o) == want and {**o} == want
    assert copy.deepcopy(o) == want and pickle.loads(pickle.dumps(o)) == want and o.copy() == want
    assert _make(5, []) == o


def test_empty_and_plain_construction():
    from marl_demandresponse_b200.environment import ObsDict

    o = ObsDict()
    assert len(o) == 0 and list(o) == [] and o == {}
    o[1] = {"x": 1}
    assert o[1] == {"x": 1} and 1 in o and len(o) == 1
