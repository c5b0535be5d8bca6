"""GPU parity tests proper: the sm_100a kernels (through the C-ABI) against the CPU oracle."""
import numpy as np
import pytest

from gym_puzzles_b200 import abi
from parity_util import rollout_compare, single_step_compare
from oracle_lib import OracleBatch

pytestmark = pytest.mark.gpu

IDS = ["MultiRobotPuzzle-v0", "MultiRobotPuzzleHeavy-v0", "MultiRobotPuzzle-v2", "MultiRobotPuzzleHeavy-v2"]


def _assert_parity(rep):
    print(rep)
    # bit-exact: contact flags, contact-point counts, feature ids, done / truncation flags
    assert rep["flag_mismatch"] == 0
    assert rep["done_mismatch"] == 0
    # float32 tolerance 1e-5 relative (parity_util.RTOL)
    assert rep["state_tol_bad"] == 0
    assert rep["obs_not_close"] == 0
    assert rep["rew_not_close"] == 0


def test_backend_is_cuda():
    assert abi.load().backend == "cuda-sm_100a"


@pytest.mark.parametrize("env_id", IDS)
def test_rollout_parity(env_id):
    N, T = 1024, 120
    h = abi.Handle(env_id, N, seed=17, max_episode_steps=50)
    rep = rollout_compare(h, env_id, N, T, seed=17, max_episode_steps=50)
    _assert_parity(rep)
    assert rep["dones"] >= 2 * N
    # the device path is expected to be bit-identical almost everywhere (only libm-level sin/cos/pow may differ)
    assert rep["state_bit_bad"] <= max(1, N // 100)
    h.close()


def test_one_step_equivalence_65536_v0():
    """BASELINE.json configs[1]: MultiRobotPuzzle-v0 batched 65,536 envs, one-step state equivalence from identical states."""
    N = 65536
    o = OracleBatch("MultiRobotPuzzle-v0", N, seed=17, nthreads=8)
    o.reset()
    for t in range(40):   # reach contact-rich states
        o.step(o.sample_actions(t))
    states = o.get_state()
    h = abi.Handle("MultiRobotPuzzle-v0", N, seed=17)
    rep = single_step_compare(h, o, states, o.sample_actions(1000))
    _assert_parity(rep)
    h.close()


def test_sharding_invariance():
    """RNG is keyed by global env id: two shards of 256 == one batch of 512."""
    a = abi.Handle("MultiRobotPuzzleHeavy-v0", 512, seed=5, max_episode_steps=40)
    b0 = abi.Handle("MultiRobotPuzzleHeavy-v0", 256, seed=5, max_episode_steps=40, env_id_base=0)
    b1 = abi.Handle("MultiRobotPuzzleHeavy-v0", 256, seed=5, max_episode_steps=40, env_id_base=256)
    oa = a.reset_host()
    ob = np.concatenate([b0.reset_host(), b1.reset_host()])
    assert np.array_equal(oa, ob)
    rng = np.random.default_rng(0)
    for t in range(100):
        act = rng.uniform(-1, 1, (512, 15)).astype(np.float32)
        ra = a.step_host(act)
        r0, r1 = b0.step_host(act[:256]), b1.step_host(act[256:])
        for x, y0, y1 in zip(ra, r0, r1):
            assert np.array_equal(x, np.concatenate([y0, y1]))
    assert np.array_equal(a.get_state(), np.concatenate([b0.get_state(), b1.get_state()]))


def test_chunked_streams_are_invariant(monkeypatch):
    """Chunk pipelines on separate streams (mrp_step) and chunked H2D / D2H overlap (mrp_step_host) give the same
    results as one chunk."""
    import torch

    N = 8192 + 300
    monkeypatch.setenv("MRP_CHUNKS", "1")
    monkeypatch.setenv("MRP_CHUNKS_HOST", "1")
    ref = abi.Handle("MultiRobotPuzzleHeavy-v0", N, seed=4, max_episode_steps=25)
    monkeypatch.setenv("MRP_CHUNKS", "4")
    monkeypatch.setenv("MRP_CHUNKS_HOST", "8")
    monkeypatch.setenv("MRP_HOST_WAVES", "1")
    chk = abi.Handle("MultiRobotPuzzleHeavy-v0", N, seed=4, max_episode_steps=25)
    # front-half waves + big islands on their own kernel: the large-batch flow of mrp_step_host (pinned buffers: early row copies)
    monkeypatch.setenv("MRP_BIG", "1")
    monkeypatch.setenv("MRP_SPARES", "1")      # ... with spare episodes: the auto-resets of the back chunks copy what the refill pass prepared
    monkeypatch.setenv("MRP_REFILL_MIN", "1")
    wavs = []
    for w in ("2", "3", "4"):
        monkeypatch.setenv("MRP_HOST_WAVES", w)
        wavs.append(abi.Handle("MultiRobotPuzzleHeavy-v0", N, seed=4, max_episode_steps=25))
    pin = (torch.empty(N, 40).pin_memory().numpy(), torch.empty(N).pin_memory().numpy(),
           torch.empty(N, dtype=torch.uint8).pin_memory().numpy(), torch.empty(N, dtype=torch.uint8).pin_memory().numpy())
    r0 = ref.reset_host()
    assert np.array_equal(r0, chk.reset_host()) and all(np.array_equal(r0, wav.reset_host()) for wav in wavs)
    rng = np.random.default_rng(2)
    for t in range(60):
        act = rng.uniform(-1, 1, (N, 15)).astype(np.float32)
        if t % 2:   # host-buffer call
            out_ref = ref.step_host(act)
            for x, y in zip(out_ref, chk.step_host(act)):
                assert np.array_equal(x, y)
            for wav in wavs:
                for x, y in zip(out_ref, wav.step_host(act, *pin) if t % 4 == 1 else wav.step_host(act)):
                    assert np.array_equal(x, y)
        else:       # device-resident call on torch's current stream
            a = torch.from_numpy(act).cuda()
            ref.step(a.data_ptr()); chk.step(a.data_ptr())
            for wav in wavs:
                wav.step(a.data_ptr())
            torch.cuda.synchronize()
    assert np.array_equal(ref.get_state(), chk.get_state()) and all(np.array_equal(ref.get_state(), wav.get_state()) for wav in wavs)
    sr, sc = ref.stats(), chk.stats()
    for k in ("episodes", "done_by_env", "truncated", "sum_length", "overflow"):
        assert sr[k] == sc[k]
    assert sr["episodes"] > 0 and abs(sr["sum_return"] - sc["sum_return"]) <= 1e-9 * abs(sr["sum_return"])   # atomics order


@pytest.mark.parametrize("device_path", [False, True])
@pytest.mark.parametrize("env_id,n_agents", [("MultiRobotPuzzle-v2", 5), ("MultiRobotPuzzleHeavy-v2", 3)])
def test_v2_more_agents_parity(env_id, n_agents, device_path):
    """BASELINE.json configs[3] with the num_agents ctor kw (mrp02:139): wide-capacity build of the kernels, through
    mrp_step_host and through the device-resident mrp_step."""
    N, T = 512, 100
    h = abi.Handle(env_id, N, seed=17, max_episode_steps=50, n_agents=n_agents)
    assert h.layout.max_contacts > 32
    rep = rollout_compare(h, env_id, N, T, seed=17, max_episode_steps=50, n_agents=n_agents, device_path=device_path)
    _assert_parity(rep)
    assert rep["dones"] >= N and h.stats()["overflow"] == 0
    h.close()


def test_v2_more_agents_overlapped_post(monkeypatch):
    """Wide build with the overlapped k_post (the default from 65,536 envs): the event pass of the task-free group uses
    the env's slice of the task pool as scratch and must not run beside the solver kernels, whose records start at the
    pool's base (round-1 ADVICE: data race).  Low env indices with many live records are the ones at risk."""
    monkeypatch.setenv("MRP_OVERLAP_POST", "1")
    env_id, n_agents, N, T = "MultiRobotPuzzle-v2", 5, 4096, 80
    h = abi.Handle(env_id, N, seed=23, max_episode_steps=40, n_agents=n_agents)
    rep = rollout_compare(h, env_id, N, T, seed=23, max_episode_steps=40, n_agents=n_agents, device_path=True, nthreads=16)
    _assert_parity(rep)
    assert rep["dones"] >= N
    h.close()


@pytest.mark.parametrize("N", [1, 37, 3001])
def test_overlapped_post_ragged_sizes(monkeypatch, N):
    """mrp_step's split k_post (envs without solver tasks beside the solver kernels, lists built by k_pre) with batch
    sizes that leave partial warps / CTAs; forced on (it is the default only from 65,536 envs)."""
    monkeypatch.setenv("MRP_OVERLAP_POST", "1")
    env_id = "MultiRobotPuzzleHeavy-v0"
    h = abi.Handle(env_id, N, seed=8, max_episode_steps=30)
    rep = rollout_compare(h, env_id, N, 70, seed=8, max_episode_steps=30, device_path=True)
    _assert_parity(rep)
    assert rep["dones"] >= 2 * N
    h.close()


@pytest.mark.parametrize("env_id,N", [("MultiRobotPuzzleHeavy-v0", 8192 + 300), ("MultiRobotPuzzleHeavy-v0", 1100),
                                      ("MultiRobotPuzzle-v0", 3001), ("MultiRobotPuzzleHeavy-v2", 3001)])
def test_host_step_early_copy_with_pinned_buffers(monkeypatch, env_id, N):
    """mrp_step_host with a pinned obs buffer copies a chunk's rows before its TOI-event / auto-reset passes and lets a
    kernel store the rows those passes rewrite straight into the host buffer: same results as the plain order (pageable
    buffers, or MRP_HOST_EARLY_COPY=0)."""
    import torch

    monkeypatch.setenv("MRP_CHUNKS_HOST", "8")
    monkeypatch.setenv("MRP_HOST_EARLY_COPY", "0")
    ref = abi.Handle(env_id, N, seed=4, max_episode_steps=25)
    monkeypatch.setenv("MRP_HOST_EARLY_COPY", "1")
    chk = abi.Handle(env_id, N, seed=4, max_episode_steps=25)
    assert np.array_equal(ref.reset_host(), chk.reset_host())
    pin = lambda *s, dt=torch.float32: torch.empty(s, dtype=dt).pin_memory().numpy()  # noqa: E731
    out = (pin(N, ref.obs_dim), pin(N), pin(N, dt=torch.uint8), pin(N, dt=torch.uint8))
    rng = np.random.default_rng(2)
    l_ref, l_chk = ref.launch_count, chk.launch_count
    dones = 0
    for t in range(80):
        act = rng.uniform(-1, 1, (N, ref.act_dim)).astype(np.float32)
        out[0][:] = np.nan   # every row must be rewritten by this step
        r = ref.step_host(act, *[np.empty_like(o) for o in out]) if t % 2 else ref.step_host(act)
        chk.step_host(act, *out)
        for x, y in zip(r, out):
            assert np.array_equal(x, y)
        dones += int(r[2].sum())
    assert dones >= 2 * N   # auto-reset rows were exercised
    assert np.array_equal(ref.get_state(), chk.get_state())
    assert (chk.launch_count - l_chk) > (ref.launch_count - l_ref)   # the row fix-up kernels ran
    ref.close(); chk.close()


@pytest.mark.parametrize("env_id", IDS)
def test_rollout_parity_device_path(env_id):
    """the same rollout parity through mrp_step (device-resident; what VectorEnv.step and bench.py's `value` run),
    large enough for the default overlapped k_post"""
    N, T = 65536 + 777, 45
    h = abi.Handle(env_id, N, seed=19, max_episode_steps=20)
    rep = rollout_compare(h, env_id, N, T, seed=19, max_episode_steps=20, device_path=True, nthreads=16)
    _assert_parity(rep)
    assert rep["dones"] >= 2 * N
    h.close()


def test_graph_replay_is_invariant(monkeypatch):
    """Small batches replay a captured CUDA graph (one cudaGraphLaunch per step instead of eleven launches): results are
    those of the plain launches, also when the caller's buffers or the handle's parameters change between steps
    (the graph is re-captured) and for pageable host buffers (never captured)."""
    import torch

    N, env_id = 700, "MultiRobotPuzzle-v2"
    rng = np.random.default_rng(3)
    acts = [rng.uniform(-1, 1, (N, 4)).astype(np.float32) for _ in range(4)]
    pinned = [torch.from_numpy(a).pin_memory().numpy() for a in acts]

    def run(graph, pin):
        monkeypatch.setenv("MRP_GRAPH", "1" if graph else "0")
        h = abi.Handle(env_id, N, seed=23, max_episode_steps=30)
        h.reset_host()
        out = []
        bufs = [(torch.empty(N, h.obs_dim).pin_memory().numpy(), torch.empty(N).pin_memory().numpy(),
                 torch.empty(N, dtype=torch.uint8).pin_memory().numpy(), torch.empty(N, dtype=torch.uint8).pin_memory().numpy())
                for _ in range(2)] if pin else [(None, None, None, None)] * 2
        for t in range(60):
            if t == 20:
                h.set_params(scaled_epsilon=30.0, puzzleComp=5000.0)   # baked into the captured kernel parameters
            src = pinned if pin else acts
            # same buffers for a while (replay), then alternating ones (re-capture)
            a, b = (src[0], bufs[0]) if t < 10 else (src[t % 4], bufs[t % 2])
            res = h.step_host(a, *b)
            out.append([np.array(x) for x in res])
        # device-resident calls: library action buffer, then a caller-owned one
        d = torch.from_numpy(acts[1]).cuda()
        for t in range(10):
            h.sample_actions(t)
            h.step()
        for t in range(10):
            h.step(d.data_ptr())
        torch.cuda.synchronize()
        st, n = h.get_state(), h.launch_count
        h.close()
        return out, st, n

    ref, st_ref, n_ref = run(False, True)
    for graph, pin in ((True, True), (True, False)):
        got, st, n = run(graph, pin)
        for x, y in zip(ref, got):
            for u, v in zip(x, y):
                assert np.array_equal(u, v)
        assert np.array_equal(st_ref, st)
        assert n == n_ref   # replays count the launches they contain


@pytest.mark.parametrize("device_path", [False, True])
@pytest.mark.parametrize("env_id", ["MultiRobotPuzzleHeavy-v0", "MultiRobotPuzzle-v2"])
def test_spare_episodes_gpu(env_id, device_path, monkeypatch):
    """Spare episodes (next episodes computed beside the step's kernels on a low-priority stream, auto-reset = copy): the
    default from 32,768 envs, forced on here; TimeLimit 20 keeps a steady stream of resets going."""
    monkeypatch.setenv("MRP_SPARES", "1")
    monkeypatch.setenv("MRP_OVERLAP_POST", "1")
    N, T = 4096, 90
    h = abi.Handle(env_id, N, seed=29, max_episode_steps=20)
    rep = rollout_compare(h, env_id, N, T, seed=29, max_episode_steps=20, device_path=device_path, nthreads=16)
    _assert_parity(rep)
    assert rep["dones"] >= 4 * N
    h.close()
