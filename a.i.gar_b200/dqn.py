"""Consumer side of BASELINE.json configs[4]: observations handed zero-copy (DLPack) to a torch DQN forward, arg-max
mapped to the reference's 5x5 action table, fed back to the env on the same stream.

Mirrors: the MLP shape of src/model/network.py:208-241 with src/model/networkParameters.py:110-118 (Q_LAYERS =
(100, 100), relu hidden, linear output, NUM_ACTIONS = 25, glorot_uniform), the discrete action table of
src/model/network.py:46-65, and the collector tick of src/aigar.py:844-849.  The GEMMs are plain library calls
(torch / cuBLAS); the env step is this repo's kernels."""
import math


def square_action_table(num_actions=25, enable_split=False, enable_eject=False):
    """createDiscreteActionsSquare (src/model/network.py:46-65) as a list of [x, y, split, eject]."""
    side = int(math.isqrt(num_actions))
    if side * side != num_actions:
        raise ValueError("Number of Actions has to be a perfect square for this mode.")
    actions = []
    for row in range(side):
        for col in range(side):
            x = col / side + 1 / side / 2
            y = row / side + 1 / side / 2
            actions.append([x, y, 0, 0])
            if enable_split:
                actions.append([x, y, 1, 0])
            if enable_eject:
                actions.append([x, y, 0, 1])
    return actions


def make_dqn(state_len, layers=(100, 100), num_actions=25, device="cuda", dtype=None, seed=0):
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    mods, prev = [], state_len
    for width in layers:
        mods += [torch.nn.Linear(prev, width), torch.nn.ReLU()]
        prev = width
    mods.append(torch.nn.Linear(prev, num_actions))
    net = torch.nn.Sequential(*mods)
    for m in net:
        if isinstance(m, torch.nn.Linear):  # INITIALIZER = "glorot_uniform" for kernels and biases' fan-in rule
            torch.nn.init.xavier_uniform_(m.weight, generator=g)
            torch.nn.init.zeros_(m.bias)
    net = net.to(device)
    if dtype is not None:
        net = net.to(dtype)
    return net.eval()


class DQNDriver(object):
    """obs (DLPack, zero copy) -> MLP -> arg-max -> action table -> env.step, all on torch's current stream."""

    def __init__(self, batch, net=None, num_actions=25, epsilon=0.0, seed=0):
        import torch
        self.torch = torch
        self.batch = batch
        L = batch.layout
        self.net = net if net is not None else make_dqn(L.state_len, num_actions=num_actions, device=batch.device, seed=seed)
        table = square_action_table(num_actions, bool(batch.cfg.enable_split), bool(batch.cfg.enable_eject))
        self.table = torch.tensor(table, dtype=torch.float32, device=batch.device)
        self.epsilon = float(epsilon)
        self.gen = torch.Generator(device=batch.device).manual_seed(seed)
        # the observation buffer seen through DLPack: same memory as batch.obs, no copy
        self.obs_view = torch.from_dlpack(batch.obs_dlpack())
        assert self.obs_view.data_ptr() == batch.obs.data_ptr()
        self.period = batch.cfg.frame_skip + 1

    @property
    def n_actions(self):
        return self.table.shape[0]

    def decide(self):
        """learningAlg.decideMove for every agent at once (src/model/qLearning.py:217-243, e-greedy)."""
        torch = self.torch
        E, A, L = self.obs_view.shape
        with torch.no_grad():
            q = self.net(self.obs_view.view(E * A, L).to(next(self.net.parameters()).dtype))
            idx = q.argmax(dim=1)
            if self.epsilon > 0:
                explore = torch.rand(E * A, device=idx.device, generator=self.gen) < self.epsilon
                rnd = torch.randint(0, self.n_actions, (E * A,), device=idx.device, generator=self.gen)
                idx = torch.where(explore, rnd, idx)
        self.last_idx = idx.view(E, A)
        return self.table[idx].view(E, A, 4)

    def tick(self):
        """One collector tick (src/aigar.py:844-849): decide on the current observation, advance FRAME_SKIP_RATE + 1
        frames, observe again — two launches of ours (none if the decision is the only consumer) plus the MLP."""
        actions = self.decide()
        self.batch.step_observe(actions, self.period)

    def run(self, n_decisions, use_graph=False):
        """n_decisions collector ticks.  use_graph: capture one tick (MLP + arg-max + table lookup + our step kernel, all on
        one stream, no host synchronisation inside) into a CUDA graph and replay it — the tick is launch-bound below ~100k
        envs."""
        torch = self.torch
        self.batch.observe()
        if not use_graph or self.epsilon > 0:
            for _ in range(n_decisions):
                self.tick()
            return
        if getattr(self, "_graph", None) is None:
            side = torch.cuda.Stream(device=self.batch.device)
            side.wait_stream(torch.cuda.current_stream(self.batch.device))
            with torch.cuda.stream(side):  # warm-up outside capture (lazy cuBLAS / attribute initialisation)
                self.tick()
                self.tick()
            torch.cuda.current_stream(self.batch.device).wait_stream(side)
            torch.cuda.synchronize(self.batch.device)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self.tick()
            n_decisions -= 2  # the two warm-up ticks advanced the envs (capturing records, it does not execute)
        for _ in range(max(n_decisions, 0)):
            self._graph.replay()
