"""The warp-cooperative Bvh::find_best of parry_b200/csrc/traverse.cuh (bvh_find_best_cost) restated lane by lane in Python and
checked against the plain per-query descent it replaces: every lane must evaluate the same node costs and the same leaves in the same
order — waiting at a leaf until the warp runs the leaf code only changes WHEN a lane's leaf is evaluated, never against which `best`.
(The CUDA kernels themselves are checked bit for bit against the oracle in the -m gpu tests; this pins the scheduling argument.)"""
import numpy as np

FMAX = float(np.finfo(np.float32).max)


def make_tree(rng, n_leaves):
    """Random binary tree in the library's layout: node = (left, right), child = (id, is_leaf); leaves carry a box [lo, hi] on a line."""
    items = [("leaf", i) for i in range(n_leaves)]
    nodes = []
    while len(items) > 1:
        i = int(rng.integers(0, len(items) - 1))
        a, b = items[i], items[i + 1]
        nodes.append((a, b))
        items[i:i + 2] = [("node", len(nodes) - 1)]
    root = items[0][1]
    lo = np.sort(rng.random(n_leaves) * 10.0)
    leaf_box = [(float(x), float(x + rng.random() * 2.0)) for x in lo]

    def box(child):
        kind, i = child
        if kind == "leaf":
            return leaf_box[i]
        l, r = box(nodes[i][0]), box(nodes[i][1])
        return (min(l[0], r[0]), max(l[1], r[1]))
    return nodes, root, box


def query_fns(q, box, leaf_value, log):
    def cost(child, bound):                       # distance from q to the child's box (a conservative bound of the leaf value)
        lo, hi = box(child)
        c = max(lo - q, q - hi, 0.0)
        log.append(("cost", child))
        return c

    def leaf(i):
        log.append(("leaf", i))
        return leaf_value(i, q)
    return cost, leaf


def sequential(nodes, root, max_cost, cost, leaf):
    """bvh_find_best / the reference's find_best with the library's tie rule (visit equal scores once something was found)."""
    best, found, best_id = max_cost, False, None
    stack, curr = [], root
    while True:
        l, r = nodes[curr]
        ls, rs = cost(l, best), cost(r, best)
        if ls > rs:
            l, r, ls, rs = r, l, rs, ls
        found_next = False
        for child, s in ((l, ls), (r, rs)):
            if s != FMAX and (s < best or (found and s == best)):
                if child[0] == "leaf":
                    d = leaf(child[1])
                    if d < best or (found and d == best and child[1] < best_id):
                        best, found, best_id = d, True, child[1]
                elif found_next:
                    stack.append(child[1])
                else:
                    curr, found_next = child[1], True
        if not found_next:
            if not stack:
                return best, best_id
            curr = stack.pop()


def cooperative(nodes, root, max_cost, lanes, min_lanes):
    """bvh_find_best_cost: one state machine per lane (VISIT / RESOLVE / WAIT / DONE), leaf code only in the warp's leaf phase."""
    VISIT, RESOLVE, WAIT, DONE = range(4)
    st = [dict(state=VISIT, curr=root, stack=[], best=max_cost, found=False, best_id=None, ent=None, pi=2, found_next=False) for _ in lanes]
    leaf_phases = 0
    while True:
        for k, (cost, leaf) in enumerate(lanes):
            s = st[k]
            if s["state"] == VISIT:
                l, r = nodes[s["curr"]]
                ls, rs = cost(l, s["best"]), cost(r, s["best"])
                if ls > rs:
                    l, r, ls, rs = r, l, rs, ls
                s["ent"], s["pi"], s["found_next"], s["state"] = ((l, ls), (r, rs)), 0, False, RESOLVE
            if s["state"] == RESOLVE:
                while s["pi"] < 2:
                    child, sc = s["ent"][s["pi"]]
                    if sc != FMAX and (sc < s["best"] or (s["found"] and sc == s["best"])):
                        if child[0] == "leaf":
                            s["state"] = WAIT
                            break
                        if s["found_next"]:
                            s["stack"].append(child[1])
                        else:
                            s["curr"], s["found_next"] = child[1], True
                    s["pi"] += 1
                if s["state"] == RESOLVE:
                    if s["found_next"]:
                        s["state"] = VISIT
                    elif s["stack"]:
                        s["curr"], s["state"] = s["stack"].pop(), VISIT
                    else:
                        s["state"] = DONE
        waiting = [k for k, s in enumerate(st) if s["state"] == WAIT]
        moving = [k for k, s in enumerate(st) if s["state"] == VISIT]
        if not waiting and not moving:
            break
        if waiting and (not moving or len(waiting) >= min_lanes):
            leaf_phases += 1
            for k in waiting:
                s = st[k]
                i = s["ent"][s["pi"]][0][1]
                d = lanes[k][1](i)
                if d < s["best"] or (s["found"] and d == s["best"] and i < s["best_id"]):
                    s["best"], s["found"], s["best_id"] = d, True, i
                s["pi"] += 1
                s["state"] = RESOLVE
    return [(s["best"], s["best_id"]) for s in st], leaf_phases


def test_cooperative_descent_evaluates_what_the_sequential_one_does():
    rng = np.random.default_rng(7)
    for trial in range(40):
        n_leaves = int(rng.integers(2, 200))
        nodes, root, box = make_tree(rng, n_leaves)
        bump = rng.random(n_leaves) * 0.5
        quant = trial % 2 == 0                   # quantised values => exact ties between leaves and between node scores

        def leaf_value(i, q, bump=bump, quant=quant, box=box):
            lo, hi = box(("leaf", i))
            d = max(lo - q, q - hi, 0.0) + float(bump[i])
            return float(np.round(d * 2) / 2) if quant else d
        queries = rng.random(32) * 14.0 - 2.0
        max_cost = FMAX if trial % 3 else 3.0
        seq_logs, seq_res = [], []
        for q in queries:
            log = []
            cost, leaf = query_fns(float(q), box, leaf_value, log)
            seq_res.append(sequential(nodes, root, max_cost, cost, leaf))
            seq_logs.append(log)
        for min_lanes in (1, 12, 32):
            logs = [[] for _ in queries]
            lanes = [query_fns(float(q), box, leaf_value, logs[k]) for k, q in enumerate(queries)]
            res, phases = cooperative(nodes, root, max_cost, lanes, min_lanes)
            assert res == seq_res
            assert logs == seq_logs              # same costs, same leaves, same order, per lane
            assert phases >= 1 or all(r[1] is None for r in res)
