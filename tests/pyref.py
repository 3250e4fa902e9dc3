"""SECOND, INDEPENDENT restatement of the reference's search, for cross-checking oracle/omok_oracle.c (test infrastructure).

The C oracle and the CUDA kernels were written from one reading of the Rust.  Nothing the reference holds pins the mcts /
alpha-zero crates (they contain no #[test]), and no Rust toolchain exists here, so the next best guard against a shared
misreading is a restatement written separately, straight from the Rust sources, in a different shape: plain Python
objects, children lists in creation order, `n` as an unbounded int (u64), `max_by` as an explicit loop, every f32
operation a numpy float32 operation (round-to-nearest after each one, no fused multiply-add).  tests/test_oracle_cross.py
runs both on the same parameters and compares trees bit for bit.

What is restated from the reference (file:line relative to /root/reference/):
  Environment            environment/src/lib.rs:62-193
  Node / MCTS            mcts/src/node.rs:10-99, mcts/src/lib.rs:13-93
  BoardState             alpha-zero/src/mcts_node.rs:7-45
  execute                alpha-zero/src/parallel_mcts_executor.rs:26-286
  Agent                  alpha-zero/src/agent.rs:10-241
What is NOT the reference's and is shared by specification with the oracle and the kernels: the counter
based random stream that replaces thread_rng() (restated here from its formula), the bounded-integer algorithm of rand
0.8.5 `UniformInt::sample_single` and `WeightedIndex` / `Uniform<f32>` (restated from the crate's published algorithm),
the exact hash evaluator (restated from its formula), and two deterministic primitives taken from the oracle library
as black boxes: the Dirichlet(alpha) sample and exp() for Boltzmann sampling.
"""
from __future__ import annotations

import numpy as np

F = np.float32
EPS = F(np.finfo(np.float32).eps)  # f32::EPSILON
SIDE, CELLS = 9, 81
EMPTY, BLACK, WHITE = 0, 1, 2
IN_PROGRESS, DRAW, BLACK_WIN, WHITE_WIN = 0, 1, 2, 3
M64 = (1 << 64) - 1
GOLDEN = 0x9E3779B97F4A7C15


# ------------------------------------------------------------------ environment/src/lib.rs
class Environment:
    def __init__(self):  # :73-79
        self.turn = 0  # Turn::Black
        self.legal_move_count = CELLS
        self.board = [EMPTY] * CELLS

    def clone(self):
        e = Environment()
        e.turn, e.legal_move_count, e.board = self.turn, self.legal_move_count, list(self.board)
        return e

    def _count_serial_stones(self, turn, index, offsets):  # :168-193
        stone = BLACK if turn == 0 else WHITE
        x0, y0 = index % SIDE, index // SIDE
        count = 0
        for ox, oy in offsets:
            x, y = x0 + ox, y0 + oy
            if x < 0 or SIDE <= x or y < 0 or SIDE <= y:
                break
            if self.board[y * SIDE + x] != stone:
                break
            count += 1
        return count

    def place_stone(self, index):  # :104-166; None == Option::None
        if self.board[index] != EMPTY:
            return None
        self.legal_move_count -= 1
        self.board[index] = BLACK if self.turn == 0 else WHITE
        rng5 = range(1, 6)
        c = self._count_serial_stones
        horizontal = 1 + c(self.turn, index, [(-k, 0) for k in rng5]) + c(self.turn, index, [(k, 0) for k in rng5])
        vertical = 1 + c(self.turn, index, [(0, -k) for k in rng5]) + c(self.turn, index, [(0, k) for k in rng5])
        lt_rb = 1 + c(self.turn, index, [(-k, -k) for k in rng5]) + c(self.turn, index, [(k, k) for k in rng5])
        lb_rt = 1 + c(self.turn, index, [(-k, k) for k in rng5]) + c(self.turn, index, [(k, -k) for k in rng5])
        turn = self.turn
        self.turn = 1 - self.turn
        if 5 in (horizontal, vertical, lt_rb, lb_rt):  # exactly SERIAL_STONE_COUNT
            return BLACK_WIN if turn == 0 else WHITE_WIN
        if self.legal_move_count == 0:
            return DRAW
        return IN_PROGRESS


# ------------------------------------------------------------------ specified primitives (not the reference's)
def _mix64(z):
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def rng_u32(seed, stream, counter):
    k = _mix64((seed + GOLDEN * (stream + 1)) & M64)
    return _mix64((k + GOLDEN * (counter + 1)) & M64) >> 32


class Stream:
    """One agent's random stream: (seed, stream id) and the position in it."""

    def __init__(self, seed, stream):
        self.seed, self.stream, self.counter, self.noise_epoch = seed, stream, 0, 0

    def next_u32(self):
        v = rng_u32(self.seed, self.stream, self.counter)
        self.counter += 1
        return v

    def below(self, bound):
        """rand 0.8.5 UniformInt::<u32>::sample_single(0, bound): widening multiply, zone = (range << lz) - 1."""
        lz = 32 - bound.bit_length()
        zone = ((bound << lz) - 1) & 0xFFFFFFFF
        while True:
            m = self.next_u32() * bound
            if (m & 0xFFFFFFFF) <= zone:
                return m >> 32


def hash_eval(env, opponent_mode):
    """The exact hash evaluator: P in (0,1], V in [-1,1) from the packed board (spec: DESIGN.md, omk_device.cuh)."""
    bw = [0] * 6
    for i, s in enumerate(env.board):
        if s == BLACK:
            bw[i >> 5] |= 1 << (i & 31)
        elif s == WHITE:
            bw[3 + (i >> 5)] |= 1 << (i & 31)
    h = 0x243F6A8885A308D3
    h = _mix64(h ^ (bw[0] | (bw[1] << 32)))
    h = _mix64(h ^ (bw[2] | (bw[3] << 32)))
    h = _mix64(h ^ (bw[4] | (bw[5] << 32)))
    h = _mix64(h ^ (env.turn | ((1 if opponent_mode else 0) << 1)))
    p = [F((_mix64((h + GOLDEN * (a + 1)) & M64) >> 40) + 1) * F(2.0 ** -24) for a in range(CELLS)]
    v = F(_mix64((h + GOLDEN * 82) & M64) >> 40) * F(2.0 ** -23) - F(1.0)
    return p, v


def _oracle_lib():
    from oracle import oracle as o  # black-box primitives only: Dirichlet sample, deterministic exp

    return o.lib()


def dirichlet81(stream: Stream, alpha):
    import ctypes as C

    out = np.zeros(CELLS, dtype=np.float32)
    _oracle_lib().orc_dirichlet81(stream.seed, stream.stream, stream.noise_epoch, float(alpha), out.ctypes.data_as(C.POINTER(C.c_float)))
    stream.noise_epoch += 1
    return [F(x) for x in out]


def det_exp_f32(x):
    return F(_oracle_lib().orc_det_exp(float(x)))


def f32_sum(values):  # Iterator::sum::<f32>(): sequential, index order
    s = F(0.0)
    for v in values:
        s = F(s + v)
    return s


def total_cmp_key(x):  # f32::total_cmp as an integer key
    b = int(np.array(x, dtype=np.float32).view(np.int32))
    return b ^ (((b >> 31) & 0xFFFFFFFF) >> 1) if b < 0 else b


# ------------------------------------------------------------------ mcts crate + BoardState
class State:  # alpha-zero/src/mcts_node.rs:7-34
    def __init__(self, env, status, policy, z):
        self.env, self.status, self.policy, self.z = env, status, policy, F(z)

    def is_terminal(self):
        return self.status != IN_PROGRESS

    def available_actions_len(self):
        return self.env.legal_move_count

    def is_available_action(self, action):
        return self.env.board[action] == EMPTY


class Node:  # mcts/src/node.rs:10-37
    def __init__(self, parent, action, p, state):
        self.parent, self.action, self.children = parent, action, []
        self.p, self.w, self.n, self.state = F(p), F(0.0), 0, state

    def select_leaf(self, selector):  # :39-59
        node = self
        while True:
            if len(node.children) != node.state.available_actions_len():
                return node
            if len(node.children) == 0:
                return node
            node = node.children[selector(node, node.children)]

    def expand(self, action, state):  # :61-81
        if any(ch.action == action for ch in self.children):
            return None
        child = Node(self, action, self.state.policy[action], state)
        self.children.append(child)
        return child

    def propagate(self, w):  # :83-99
        node, w = self, F(w)
        while True:
            node.n += 1
            node.w = F(node.w + w)
            w = F(-w)
            if node.parent is None:
                break
            node = node.parent


class MCTS:  # mcts/src/lib.rs:13-78
    def __init__(self, root_state):
        self.root = Node(None, None, 1.0, root_state)

    def transition(self, children_index):  # :47-78 (the freed siblings simply become unreachable)
        new_root = self.root.children[children_index]
        new_root.parent = None
        new_root.n = sum(ch.n for ch in new_root.children)
        self.root = new_root

    def node_count(self):
        stack, k = [self.root], 0
        while stack:
            nd = stack.pop()
            k += 1
            stack.extend(nd.children)
        return k


def compute_ucb_1(parent_n, node, c):  # parallel_mcts_executor.rs:277-286
    n = node.n
    q = F(node.w / F(F(n) + EPS))
    bias = F(np.sqrt(F(parent_n)) / F(1 + n))
    return F(q + F(F(F(c) * node.p) * bias))


# ------------------------------------------------------------------ alpha-zero/src/agent.rs
class Agent:
    def __init__(self, evaluate_p, seed, stream):  # Agent::new :16-35
        self.env = Environment()
        policy = list(evaluate_p(self.env, False))  # raw, unmasked
        self.mcts = MCTS(State(self.env.clone(), IN_PROGRESS, policy, 0.0))
        self.rng = Stream(seed, stream)

    def compute_policy(self):  # :43-77
        root = self.mcts.root
        if not root.children:
            return None
        s = F(0.0)
        policy = [F(0.0)] * CELLS
        for ch in root.children:
            n = F(ch.n)
            s = F(s + n)
            policy[ch.action] = n
        if s < EPS:
            return None
        inv = F(F(1.0) / s)
        return [F(v * inv) for v in policy]

    def sample_action(self, boltzmann_temperature=None):  # :83-137
        policy = self.compute_policy()
        if policy is None:
            return None
        if boltzmann_temperature is None:  # Best: max_by(total_cmp) keeps the LAST maximum
            best, best_key = 0, total_cmp_key(policy[0])
            for i in range(1, CELLS):
                k = total_cmp_key(policy[i])
                if k >= best_key:
                    best, best_key = i, k
            return best, policy
        s = F(0.0)
        heated = [F(0.0)] * CELLS
        tinv = F(F(1.0) / F(boltzmann_temperature))
        for a in range(CELLS):
            if policy[a] < EPS:
                continue
            h = det_exp_f32(F(policy[a] * tinv))
            s = F(s + h)
            heated[a] = h
        inv = F(F(1.0) / s)
        heated = [F(h * inv) for h in heated]
        # rand 0.8.5 WeightedIndex::new + sample: running sums of all but the last weight, Uniform<f32>::new(0, total)
        cum, total = [], heated[0]
        for w in heated[1:]:
            cum.append(total)
            total = F(total + w)
        scale = total
        max_rand = F(1.0) - F(2.0 ** -23)
        while F(F(scale * max_rand) + F(0.0)) >= total:
            scale = np.array(np.array(scale, np.float32).view(np.uint32) - np.uint32(1), np.uint32).view(np.float32)[()]
        r = self.rng.next_u32()
        value1_2 = np.array((r >> 9) | 0x3F800000, np.uint32).view(np.float32)[()]
        chosen = F(F(F(value1_2 - F(1.0)) * scale) + F(0.0))
        idx = 0
        while idx < CELLS - 1 and cum[idx] <= chosen:  # partition_point(|w| w <= chosen)
            idx += 1
        return idx, policy

    def ensure_action_exists(self, action, evaluate_p):  # :144-197
        if CELLS <= action:
            return
        env = self.env.clone()
        env.place_stone(action)  # result ignored
        policy = list(evaluate_p(env, True))  # EnvTurnMode::Opponent
        policy[action] = F(0.0)
        for a in range(CELLS):
            if not self.mcts.root.state.is_available_action(a):
                policy[a] = F(0.0)
        s = f32_sum(policy)
        if EPS <= s:
            inv = F(F(1.0) / s)
            policy = [F(v * inv) for v in policy]
        self.mcts.root.expand(action, State(env, IN_PROGRESS, policy, 0.0))

    def play_action(self, action):  # :206-232
        if self.mcts.root.state.is_terminal():
            return None
        index = next((i for i, ch in enumerate(self.mcts.root.children) if ch.action == action), None)
        if index is None:
            return None
        status = self.env.place_stone(action)
        if status is None:
            return None
        self.mcts.transition(index)
        return status


# ------------------------------------------------------------------ alpha-zero/src/parallel_mcts_executor.rs:26-270
C_PUCT = 1.0


def _selector(parent, children):  # :81-90
    parent_n = max(1, parent.n)
    best, best_key = 0, None
    for i, ch in enumerate(children):
        k = total_cmp_key(compute_ucb_1(parent_n, ch, C_PUCT))
        if best_key is None or k >= best_key:  # max_by returns the last of equal maxima
            best, best_key = i, k
    return best


def _generate_requests(agent: Agent, first_round, batch_size, epsilon, alpha):
    root = agent.mcts.root
    if first_round:  # :48-76 (runs with epsilon == 0 too: a pure renormalisation)
        eps = F(epsilon)
        noise = dirichlet81(agent.rng, alpha) if epsilon != 0.0 else [F(0.0)] * CELLS
        policy = root.state.policy
        for a in range(CELLS):
            policy[a] = F(F(F(F(1.0) - eps) * policy[a]) + F(eps * noise[a]))
        inv = F(F(1.0) / f32_sum(policy))  # no EPSILON guard
        for a in range(CELLS):
            policy[a] = F(policy[a] * inv)
        for ch in root.children:
            ch.p = policy[ch.action]
    requests = []
    for _ in range(batch_size):
        node = root.select_leaf(_selector)
        if node.state.is_terminal():  # :92-97
            node.propagate(node.state.z)
            continue
        taken = {ch.action for ch in node.children}
        available = [a for a in range(CELLS) if node.state.is_available_action(a) and a not in taken]  # :101-116
        if not available:
            continue
        action = available[agent.rng.below(len(available))]  # SliceRandom::choose
        env = node.state.env.clone()
        status = env.place_stone(action)
        reward = None if status == IN_PROGRESS else (F(0.0) if status == DRAW else F(1.0))  # :130-135
        policy = [F(1.0)] * CELLS  # :140-156 dummy uniform policy over the child's empty cells
        for a in range(CELLS):
            if env.board[a] != EMPTY:
                policy[a] = F(0.0)
        s = f32_sum(policy)
        if EPS <= s:
            inv = F(F(1.0) / s)
            policy = [F(v * inv) for v in policy]
        child = node.expand(action, State(env, status, policy, reward if reward is not None else 0.0))
        if child is None:
            continue
        if reward is not None:
            child.propagate(reward)  # :178-181
        else:
            requests.append(child)
    return requests


def execute(agents, count, batch_size, epsilon, alpha, evaluate_pv):
    """evaluate_pv(list of envs) -> (list of policies, list of values), EnvTurnMode::Player encoding."""
    processed = 0
    sims = 0
    while processed < count:
        requests = []
        for agent in agents:  # agents in index order; flat_map().collect() keeps it (:194-205)
            requests.extend(_generate_requests(agent, processed == 0, batch_size, epsilon, alpha))
        processed += batch_size
        sims += batch_size * len(agents)
        if not requests:
            continue
        policies, values = evaluate_pv([r.state.env for r in requests])
        for node, raw, val in zip(requests, policies, values):  # :222-265
            value = F(-F(val))
            policy = [F(v) for v in raw]
            for a in range(CELLS):
                if not node.state.is_available_action(a):
                    policy[a] = F(0.0)
            s = f32_sum(policy)
            if EPS <= s:
                inv = F(F(1.0) / s)
                policy = [F(v * inv) for v in policy]
            node.state.policy = policy
            for ch in node.children:
                ch.p = policy[ch.action]
            node.propagate(value)
    return sims


def hash_evaluate_p(env, opponent_mode):
    return hash_eval(env, opponent_mode)[0]


def hash_evaluate_pv(envs):
    out = [hash_eval(e, False) for e in envs]
    return [o[0] for o in out], [o[1] for o in out]
