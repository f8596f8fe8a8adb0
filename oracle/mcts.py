"""Single-threaded restatement of dream-go's tree search (`src/libdg_mcts`), SURVEY.md section 8c-3.

TEST INFRASTRUCTURE ONLY (tests/, the soak / fuzz checkers under tools/ and bench.py's CPU-baseline leg) -- never imported by the product.

Follows the reference function by function with its dense per-node tables (`BigChildrenImpl`, tree.rs:540-620):
`Node::select` (tree.rs:1311-1385) incl. the blocked argmax of asm/argmax.rs:23-76, `probe` (:1421-1471),
`insert` + `UCT::update` (:1482-1510, :125-159), `undo` (:1397-1409), `time_control::is_done`
(time_control/mod.rs:47-97), `Node::best` / `compare_children` (:1232-1283, :1524-1560), `choose` (choose.rs),
`full_forward` + `predict` (lib.rs:83-200), `create_initial_policy` / `add_valid_candidates` / `normalize_policy`
(pool/policy_helper.rs).  All arithmetic is numpy float32 in the reference's operation order.

Everything random in the reference (`thread_rng`) is an argument here: the Dirichlet sample, the symmetry of each
leaf, the uniform number of the stochastic move choice.  The pool of racing workers is replaced by its sequential
schedule: `probes_per_round` probes, then their inserts in the same order (1 = one probe in flight).

Parity status: the reference pins only invariants for this code (tree.rs:1758-1946, lib.rs:245-281); they are
asserted on this restatement in tests/test_oracle_mcts.py.  Visit counts against the reference BINARY are
undefined (its schedule depends on thread timing), see SURVEY.md section 0.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import go

F = np.float32
NEG_INF = F(-np.inf)
VLOSS_CNT = 32            # config.rs:187
MIN_LCB_VISITS = 80       # tree.rs:34

UCT_EXP = [(0, 0.77392), (800, 1.05439), (1600, 1.22798), (3200, 0.813532), (6400, 0.764326)]        # config.rs:189-190
FPU_REDUCE = [(0, 0.631571), (800, 0.431547), (1600, 0.656083), (3200, 0.429231), (6400, 0.514494)]  # config.rs:181-182
CRITICAL_VALUE = [(0, 1.91753), (800, 1.86478), (1600, 1.86943), (3200, 2.20033), (6400, 1.78053)]   # config.rs:193-194


def get_intp_value(points, x: int) -> np.float32:      # config.rs:297-312
    for i, (px, _) in enumerate(points):
        if px >= x:
            x0 = points[i - 1] if i > 0 else points[0]
            x1 = points[i]
            a = F(0.5) if x0[0] >= x1[0] else F(F(x - x0[0]) / F(x1[0] - x0[0]))
            return F(F(F(1.0) - a) * F(x0[1]) + F(a * F(x1[1])))
    return F(points[-1][1])


def argmax_f32(values: np.ndarray) -> Optional[int]:   # asm/argmax.rs:23-76 (368 entries, blocks of 8)
    so_far = NEG_INF
    index = 0
    for i in range(len(values) // 8):
        block = values[8 * i:8 * i + 8]
        m = block.max()
        if m > so_far:
            so_far = m
        eq = np.flatnonzero(block == so_far)
        if len(eq):
            index = 8 * i + int(eq[0])
    return index if np.isfinite(values[index]) else None


class Node:                                            # tree.rs:1027-1089
    def __init__(self, to_move: int, value, prior: np.ndarray):
        self.to_move = to_move
        self.initial_value = F(value)
        self.pass_count = 0
        self.total_count = 0
        self.vtotal_count = 0
        self.prior = np.full(368, NEG_INF, F)
        self.prior[:362] = np.asarray(prior[:362], F)
        self.count = np.zeros(368, np.int64)
        self.vcount = np.zeros(368, np.int64)
        self.value = np.full(368, self.initial_value, F)
        self.value_s = np.zeros(368, F)
        self.expanding = np.zeros(362, bool)
        self.ptr: List[Optional["Node"]] = [None] * 362
        self.touched: List[int] = []                   # order in which child slots were created (SmallChildrenImpl)

    def touch(self, i: int) -> None:
        if i not in self.touched:
            self.touched.append(i)

    def disqualify(self, i: int) -> None:              # tree.rs:1296-1301
        self.touch(i)
        self.value[i] = NEG_INF
        self.count[i] = 0

    def nonzero(self) -> List[int]:                    # tree.rs:871-890: slot order while sparse, index order when dense
        order = self.touched if len(self.touched) <= 8 else sorted(self.touched)
        return [i for i in order if self.count[i] != 0]

    def select(self, apply_fpu: bool):                 # tree.rs:1311-1385
        value = self.value.copy()
        n = self.total_count + self.vtotal_count
        total = self.count + self.vcount
        if apply_fpu:                                  # FPU::apply (tree.rs:190-262)
            fpu = get_intp_value(FPU_REDUCE, n)
            zero = total[:362] == 0
            reduced = np.maximum((value[:362] - fpu).astype(F), F(0.0))
            reduced = np.where(np.isnan(reduced), F(0.0), reduced)
            value[:362] = np.where(zero, reduced, value[:362])
        value[362:] = NEG_INF
        sqrt_n = F(math.sqrt(F(1 + n)))                # UCT::get_impl (tree.rs:69-121)
        u = F(get_intp_value(UCT_EXP, n) * sqrt_n)
        with np.errstate(invalid="ignore"):
            bonus = np.where(total == 0, u, (u / (1 + total).astype(F)).astype(F)).astype(F)
            value = (value + (self.prior * bonus).astype(F)).astype(F)
        i = argmax_f32(value)
        if i is None:
            return "noresult", None
        self.touch(i)
        was_expanding = bool(self.expanding[i])
        self.expanding[i] = True
        if was_expanding and self.ptr[i] is None:
            return "conflict", None
        self.vcount[i] += VLOSS_CNT
        self.vtotal_count += VLOSS_CNT
        return "found", i


Trace = List[Tuple[Node, int]]


def undo(trace: Trace, undo_expanding: bool) -> None:  # tree.rs:1397-1409
    for node, i in trace:
        node.vtotal_count -= VLOSS_CNT
        node.vcount[i] -= VLOSS_CNT
        if undo_expanding and node.ptr[i] is None:
            node.expanding[i] = False


def probe(root: Node, board: "go.Board"):              # tree.rs:1421-1471
    trace: Trace = []
    current = root
    while True:
        status, i = current.select(apply_fpu=len(trace) > 0)
        if status == "conflict":
            undo(trace, False)
            return "conflict", None
        if status == "noresult":
            return "noresult", None
        trace.append((current, i))
        if i != 361:
            board.place_index(current.to_move, i)
        elif current.pass_count >= 1:
            break
        child = current.ptr[i]
        if child is None:
            break
        current = child
    return "found", trace


def insert(trace: Trace, color: int, value, prior: np.ndarray) -> None:   # tree.rs:1482-1510
    value = F(value)
    if trace:
        node, i = trace[-1]
        if node.ptr[i] is None:
            nxt = Node(color, value, prior)
            if i == 361:
                nxt.pass_count = node.pass_count + 1
            node.ptr[i] = nxt
    for node, i in trace:                              # UCT::update (tree.rs:125-159)
        v = value if color == node.to_move else F(F(1.0) - value)
        node.total_count += 1
        node.vtotal_count -= VLOSS_CNT
        prev, prev_s = node.value[i], node.value_s[i]
        prev_count = int(node.count[i])
        node.count[i] = prev_count + 1
        nxt = F(prev + F(F(v - prev) / F(prev_count + 1)))
        node.value[i] = nxt
        node.value_s[i] = F(prev_s + F(F(v - prev) * F(v - nxt)))
        node.vcount[i] -= VLOSS_CNT


def is_done(root: Node, limit: int) -> bool:           # time_control/mod.rs:83-97 with RolloutLimit
    if root.total_count == 0:
        return False
    if root.total_count >= limit:
        return True
    remaining = limit - root.total_count
    counts = root.count[:362]
    top_1 = int(np.argmax(counts))                     # min_promote_rollouts (:47-74); ties do not change the difference
    top_2 = 1 if top_1 == 0 else 0
    for i in root.nonzero():
        if i != top_1 and counts[i] > counts[top_2]:
            top_2 = i
    c1, c2 = int(counts[top_1]), int(counts[top_2])
    return (c1 - c2 if c1 > c2 else 0) > remaining


def normal_lcb_m(p_hat, p_std, n: int, m: int):        # libdg_utils/lcb.rs:28-36
    if n > 0:
        z = get_intp_value(CRITICAL_VALUE, m)
        return F(p_hat - F(F(z * p_std) / F(math.sqrt(F(n)))))
    return F(0.0)


def compare_children(node: Node, a: int, b: int) -> int:   # tree.rs:1524-1560
    def cmp(x, y):
        return -1 if x < y else (1 if x > y else 0)
    ac, bc = int(node.count[a]), int(node.count[b])
    if ac >= MIN_LCB_VISITS and bc >= MIN_LCB_VISITS:
        a_std = F(math.sqrt(F(node.value_s[a] / F(F(ac) + F(1e-5)))))
        b_std = F(math.sqrt(F(node.value_s[b] / F(F(bc) + F(1e-5)))))
        al = normal_lcb_m(node.value[a], a_std, ac, node.total_count)
        bl = normal_lcb_m(node.value[b], b_std, bc, node.total_count)
        if al != bl:
            return cmp(al, bl)
    if ac != bc:
        return cmp(ac, bc)
    if node.prior[a] != node.prior[b]:
        return cmp(node.prior[a], node.prior[b])
    return cmp(node.value[a], node.value[b])


def choose(items: Sequence[float], cutoff_percentile: float, temperature: float, at: float) -> Optional[int]:   # choose.rs:25-99
    items = [float(x) for x in items]
    total = sum(x for x in items if math.isfinite(x))
    max_value = total * (1.0 - cutoff_percentile)
    so_far, threshold = 0.0, None
    for x in sorted(items, reverse=True):
        so_far += x
        if so_far >= max_value:
            threshold = x
            break
    if threshold is None:
        return None
    cum_total = 0.0
    cum = [math.nan] * len(items)
    for i, x in enumerate(items):
        if x >= threshold:
            ratio = x / so_far if so_far != 0.0 else (math.nan if x == 0.0 else math.copysign(math.inf, x))
            cum_total += math.pow(ratio, temperature)
            cum[i] = cum_total
    target = at * cum_total
    for i, c in enumerate(cum):
        if c >= target:
            return i
    return None


def best(node: Node, temperature: float, at: float):   # tree.rs:1232-1263
    if temperature <= 9e-2:
        pick = None
        for i in node.nonzero():                       # Iterator::max_by keeps the last of equal maxima
            if pick is None or compare_children(node, pick, i) <= 0:
                pick = i
        if pick is None:
            pick = 361
        return node.value[pick], pick
    visits = [float(node.count[i]) for i in range(362)]
    i = choose(visits, 0.5, 1.0 / float(F(temperature)), at)
    if i is None:
        return node.initial_value, 361
    return node.value[i], i


def softmax(node: Node) -> np.ndarray:                  # tree.rs:1293-1306: the visit distribution the record's P[] holds
    out = np.zeros(362, F)
    total = F(0.0)
    for i in node.nonzero():
        total = F(total + F(node.count[i]))
    for i in node.nonzero():
        out[i] = F(F(node.count[i]) / total)
    return out


def forward(node: Node, index: int) -> Optional[Node]:  # tree.rs:1198-1225
    nxt = node.ptr[index]
    if nxt is None:
        if index == 361:
            nxt = Node(3 - node.to_move, 0.5, np.zeros(362, F))
            nxt.pass_count = node.pass_count + 1
            return nxt
        return None
    node.ptr[index] = None
    return nxt


# ---- pool/policy_helper.rs ------------------------------------------------------------------------------------------

def create_initial_policy(board: "go.Board", to_move: int, search: int):   # :28-75
    policy = np.full(368, NEG_INF, F)
    cand = board.policy_candidates(to_move, search)
    policy[:362][cand != 0] = 0.0
    syms = [t for t in range(8) if board.is_symmetric(t)]
    indices = np.zeros(362, np.int64)
    indices[361] = 361
    for i in range(361):
        target = min(go.symmetry_apply(t, i) for t in syms)
        indices[i] = target
        if i != target:
            policy[i] = NEG_INF
    return policy, indices


def add_valid_candidates(dst: np.ndarray, src: np.ndarray, indices: np.ndarray, transform: int) -> None:   # :87-104
    dst[361] = F(dst[361] + src[361])
    inv = go.lib().dgo_symmetry_inverse(transform)
    for i in range(361):
        j = indices[go.symmetry_apply(inv, i)]
        dst[j] = F(dst[j] + src[i])


def normalize_policy(policy: np.ndarray, sum_to) -> None:   # :113-134 (asm/sum_finite.rs:23-57, normalize_finite.rs:23-40)
    lanes = np.zeros(8, F)
    for i in range(368):
        if np.isfinite(policy[i]):
            lanes[i & 7] = F(lanes[i & 7] + policy[i])
    total = F(F(F(lanes[0] + lanes[1]) + F(lanes[2] + lanes[3])) + F(F(lanes[4] + lanes[5]) + F(lanes[6] + lanes[7])))
    finite = np.isfinite(policy)
    if total < 1e-6:
        if finite.any():
            policy[finite] = F(F(sum_to) / F(finite.sum()))
    else:
        recip = F(F(1.0) / F(total / F(sum_to)))       # the reference uses the 12-bit rcpps estimate here
        policy[:] = (policy * recip).astype(F)


class Cache:
    """`LruCache<BoardTuple, Prediction>` as `NnPredictor` uses it (predictors/nn.rs:29-82, lru_cache.rs:96-150):
    key (zobrist hash, to_move); `get` makes the entry most recent; `insert` of a present key does nothing; the least
    recent entry is dropped beyond `capacity`.  Entries are stored through `Prediction::with_transform`
    (predictor.rs:30-44) in identity orientation."""

    def __init__(self, capacity: int = 200_000):
        from collections import OrderedDict
        self.capacity = capacity
        self.entries = OrderedDict()          # last = most recent
        self.hits = 0
        self.misses = 0

    @staticmethod
    def with_transform(policy: np.ndarray, transform: int) -> np.ndarray:
        out = np.zeros(362, np.float16)
        for i in range(361):
            out[go.symmetry_apply(transform, i)] = policy[i]
        out[361] = policy[361]
        return out

    def fetch(self, board: "go.Board", to_move: int, symmetry: int):      # NnPredictor::fetch
        key = (board.zobrist_hash(), to_move)
        if key not in self.entries:
            self.misses += 1
            return None
        self.hits += 1
        self.entries.move_to_end(key)
        value, policy = self.entries[key]
        return value, self.with_transform(policy, symmetry)

    def cache(self, board: "go.Board", to_move: int, symmetry: int, value, policy: np.ndarray) -> None:   # NnPredictor::cache
        key = (board.zobrist_hash(), to_move)
        if self.capacity == 0 or key in self.entries:
            return
        inv = go.lib().dgo_symmetry_inverse(symmetry)
        self.entries[key] = (np.float16(value), self.with_transform(np.asarray(policy, np.float16), inv))
        if len(self.entries) > self.capacity:
            self.entries.popitem(last=False)


Predictor = Callable[[np.ndarray], Tuple[np.ndarray, np.ndarray]]   # features [n,361,32] f16 -> (value [n] f16, policy [n,362] f16)


def full_forward(predictor: Predictor, search: int, board: "go.Board", to_move: int, cache: Optional[Cache] = None):   # lib.rs:83-133
    """Returns (value, policy, evaluated positions)."""
    initial, indices = create_initial_policy(board, to_move, search)
    policy = initial.copy()
    value = F(0.0)
    responses, missing = {}, []
    for t in range(8):
        hit = cache.fetch(board, to_move, t) if cache is not None else None
        if hit is not None:
            responses[t] = hit
        else:
            missing.append(t)

    def add(t, v, p):
        nonlocal value, policy
        new_policy = initial.copy()
        add_valid_candidates(new_policy, np.asarray(p).astype(F), indices, t)
        normalize_policy(new_policy, 0.125)
        winrate = F(F(0.5) * F(v) + F(0.5))
        value = F(value + F(winrate * F(0.125)))
        policy[:362] = (policy[:362] + new_policy[:362]).astype(F)

    for t in range(8):                       # cached symmetries first (lib.rs:97-111) ...
        if t in responses:
            add(t, *responses[t])
    if missing:                              # ... then one batch for the rest (lib.rs:114-131)
        feats = np.stack([board.features(to_move, t) for t in missing])
        values, policies = predictor(feats)
        for k, t in enumerate(missing):
            add(t, values[k], policies[k])
            if cache is not None:
                cache.cache(board, to_move, t, values[k], policies[k])
    return value, policy, len(missing)


def dirichlet_mix(x: np.ndarray, eta: np.ndarray, beta) -> None:   # dirichlet.rs:70-75 with g/g_sum = eta supplied
    beta = F(beta)
    for i in range(len(x)):
        if np.isfinite(x[i]):
            x[i] = F(F(F(1.0) - beta) * x[i]) + F(beta * F(eta[i]))


def predict(predictor: Predictor, board: "go.Board", color: int, *, search: int = 0, deterministic: bool = False,
            num_rollout: int = 800, probes_per_round: int = 1, starting_tree: Optional[Node] = None,
            noise: Optional[np.ndarray] = None, dirichlet_noise: float = 0.25, temperature: float = 0.8,
            leaf_symmetries: Sequence[int] = (0,), choose_at: float = 0.0, cache: Optional[Cache] = None, rng=None,
            dirichlet_shape: float = 0.03):
    """`dg_mcts::predict` (lib.rs:145-200) + the worker loop of pool/worker_thread.rs in its sequential schedule.
    Returns (value, index, root, evals)."""
    value, policy, evals = full_forward(predictor, search, board, color, cache)
    if not deterministic:
        if noise is None:
            noise = rng.dirichlet(policy, float(F(dirichlet_shape)))       # drawn here, as the product's search does
        dirichlet_mix(policy[:362], noise, dirichlet_noise)
    if starting_tree is not None:
        assert starting_tree.to_move == color
        starting_tree.prior[:362] = policy[:362]
        root = starting_tree
    else:
        root = Node(color, value, policy)
    leaf = 0
    while True:
        pending = []
        hits_this_round = 0
        while len(pending) < probes_per_round:
            if is_done(root, num_rollout):
                break
            b = board.clone()
            status, trace = probe(root, b)
            if status != "found":
                break
            to_move = 3 - trace[-1][0].to_move
            sym = rng.below(8) if rng is not None else leaf_symmetries[leaf % len(leaf_symmetries)]
            leaf += 1
            hit = cache.fetch(b, to_move, sym) if cache is not None else None      # Event::predict (pool/event.rs:50-52)
            if hit is not None:
                prior, indices = create_initial_policy(b, to_move, search)
                add_valid_candidates(prior, hit[1].astype(F), indices, sym)
                normalize_policy(prior, 1.0)
                insert(trace, to_move, F(F(0.5) * F(hit[0]) + F(0.5)), prior)
                hits_this_round += 1
                if hits_this_round > 4 * probes_per_round:
                    break
                continue
            pending.append((trace, b, to_move, sym, b.features(to_move, sym)))
        if not pending:
            if hits_this_round > 0 and not is_done(root, num_rollout):
                continue
            break
        values, policies = predictor(np.stack([p[4] for p in pending]))
        evals += len(pending)
        for k, (trace, b, to_move, sym, _) in enumerate(pending):   # EventKind::Insert (worker_thread.rs:88-98)
            prior, indices = create_initial_policy(b, to_move, search)
            add_valid_candidates(prior, policies[k].astype(F), indices, sym)
            normalize_policy(prior, 1.0)
            insert(trace, to_move, F(F(0.5) * F(values[k]) + F(0.5)), prior)
            if cache is not None:
                cache.cache(b, to_move, sym, values[k], policies[k])
    t = temperature if (not deterministic and board.count() < 8) else 0.0
    if rng is not None and t > 9e-2:
        choose_at = rng.uniform()
    v, index = best(root, t, choose_at)
    return v, index, root, evals


# ---- self_play.rs:217-459 -----------------------------------------------------------------------------------------------

class Player:                                          # self_play.rs:217-241
    def __init__(self, color: int):
        self.winrate = F(0.5)                          # MovingAverage(0.5, MOMENTUM = 0.2)
        self.root: Optional[Node] = None
        self.color = color

    def num_rollout(self, max_rollout: int) -> int:
        m = F(F(F(4.0) * self.winrate) * F(F(1.0) - self.winrate))
        m = F(0.1) if m < 0.1 else m
        return int(F(m * F(max_rollout)))

    def update(self, value) -> None:
        self.winrate = F(self.winrate - F(F(0.2) * F(self.winrate - F(value))))


def get_random_komi(rng) -> float:                     # lib.rs:210-224
    v = F(rng.uniform())
    if v < F(0.4):
        return 7.5
    if v < F(0.8):
        return 6.5
    if v < F(0.9):
        return 0.5
    return float(rng.below(16) - 8) + 0.5


def self_play_one(predictor: Predictor, game_rng, *, num_rollout: int = 800, probes_per_round: int = 1, max_plies: int = 722,
                  dirichlet_noise: float = 0.25, temperature: float = 0.8, cache: Optional[Cache] = None,
                  ex_it: bool = False, num_ex_it_rollout: int = 800, recorded: Optional[list] = None):
    """`self_play_one` (self_play.rs:423-459): returns (komi, [(color, index)]).  With `ex_it` (`--ex-it`, self_play.rs:287-319,
    341-346, 384-389) a move whose value lies in [-0.8, 0.8] is, with probability 0.05, searched a second time from scratch with
    `num_ex_it_rollout` rollouts and the RECORD takes its statistics from that tree while the game goes on with the first
    search's move, value and sub-tree.  `recorded` (optional list) receives per move the rollouts of the recorded tree
    (the record's TV[] property; None where the record has none) and its visit distribution."""
    from .rng import Rng
    komi = get_random_komi(game_rng)
    board = go.Board(komi)
    players = [Player(go.BLACK), Player(go.WHITE)]
    pass_count = 0
    played = []

    def is_good_candidate(value) -> bool:              # self_play.rs:312-316 (one uniform f32 draw, only when the value qualifies)
        return bool(ex_it and F(value) >= F(-0.80) and F(value) <= F(0.80) and F(game_rng.uniform()) < F(0.05))

    def expert_iteration(allow_pass):                  # Player::ex_it (self_play.rs:287-304): self.root is None here
        v2, _, deep, _ = predict(predictor, board, players[0].color, search=0 if allow_pass else 1, deterministic=not allow_pass,
                                 num_rollout=num_ex_it_rollout, probes_per_round=probes_per_round, starting_tree=None,
                                 dirichlet_noise=dirichlet_noise, temperature=temperature, cache=cache, rng=Rng(game_rng.next()))
        return deep

    def note(tree):
        if recorded is not None:
            recorded.append(None if tree is None else (int(tree.total_count), softmax(tree)))
    while board.count() < max_plies:
        me = players[0]
        allow_pass = board.is_scorable()
        rollouts = me.num_rollout(num_rollout)
        search_rng = Rng(game_rng.next())
        tree = me.root
        me.root = None
        if rollouts > 1:
            if tree is not None and not allow_pass:
                tree.disqualify(361)                   # predict_aux (self_play.rs:252-256)
            value, index, tree, _ = predict(predictor, board, me.color, search=0 if allow_pass else 1, deterministic=not allow_pass,
                                            num_rollout=rollouts, probes_per_round=probes_per_round, starting_tree=tree,
                                            dirichlet_noise=dirichlet_noise, temperature=temperature, cache=cache, rng=search_rng)
            if not np.isfinite(value):
                index, tree = 361, None
                note(None)
            else:
                note(expert_iteration(allow_pass) if is_good_candidate(value) else tree)
                me.update(value)
                tree = forward(tree, index)
            me.root = tree
        else:                                          # self_play.rs:360-396: play from the averaged policy
            value, policy, _ = full_forward(predictor, 0 if allow_pass else 1, board, me.color, cache)
            if not allow_pass:
                policy[361] = NEG_INF
            pick = choose([float(x) for x in policy[:362]], 0.5, 1.0 / float(F(temperature)), search_rng.uniform())
            index = 361 if pick is None else pick
            note(expert_iteration(allow_pass) if is_good_candidate(value) else None)
            me.update(value)
            me.root = forward(tree, index) if tree is not None else None
        played.append((me.color, index))
        if index == 361:
            pass_count += 1
            if pass_count >= 2 and board.is_scorable():
                break
        else:
            pass_count = 0
            board.place_index(me.color, index)
        other = players[1]
        if other.root is not None:
            other.root = forward(other.root, index)
        players.reverse()
    return komi, played
