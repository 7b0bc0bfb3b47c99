"""N > 1 host logic on CPU: world_size-2 (and 4) gloo process groups exercise the z-slab partition arithmetic the
NCCL path uses (owned ranges, local lattices, halo planes) without a GPU.  Each rank derives its layout from
xsb_slab_layout, ranks exchange them over gloo and verify that they fit together; a halo exchange with the
library's plane arithmetic is then replayed on host arrays with gloo send/recv and checked against the global vector."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import exsaddle_b200 as X


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, mx, my, mz, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        lay = X.slab_layout(3, mx, my, mz, world, rank)
        lays = [None] * world
        dist.all_gather_object(lays, lay)
        NX, NY, PX, PY = 2 * mx + 1, 2 * my + 1, mx + 1, my + 1
        pu, pp = 3 * NX * NY, PX * PY
        nu_g, np_g = pu * (2 * mz + 1), pp * (mz + 1)
        # 1. owned ranges tile the global vectors, in order
        assert lays[0]["u_glob0"] == 0 and lays[0]["p_glob0"] == 0
        for a, b in zip(lays[:-1], lays[1:]):
            assert a["u_glob0"] + a["u_len"] == b["u_glob0"] and a["p_glob0"] + a["p_len"] == b["p_glob0"]
            assert a["k1"] == b["k0"]
        assert lays[-1]["u_glob0"] + lays[-1]["u_len"] == nu_g and lays[-1]["p_glob0"] + lays[-1]["p_len"] == np_g
        # 2. local lattice = layers [k0-2, k1+1) clipped; owned offsets consistent with it
        assert lay["e0"] == max(0, lay["k0"] - 2) and lay["e1"] == min(mz, lay["k1"] + 1)
        assert lay["u_off"] == (2 * (lay["k0"] - lay["e0"])) * pu and lay["u_glob0"] == 2 * lay["k0"] * pu
        nu_loc = pu * (2 * (lay["e1"] - lay["e0"]) + 1)
        assert lay["p_off"] == nu_loc + (lay["k0"] - lay["e0"]) * pp
        # 3. replay the halo exchange of comm_halo_u / comm_halo_p (2 planes up, 1 plane down for u; 1/1 for p)
        xg = np.sin(0.37 * np.arange(nu_g + np_g)) + 0.1
        loc = np.full(nu_loc + pp * (lay["e1"] - lay["e0"] + 1), np.nan)
        loc[lay["u_off"]:lay["u_off"] + lay["u_len"]] = xg[lay["u_glob0"]:lay["u_glob0"] + lay["u_len"]]
        loc[lay["p_off"]:lay["p_off"] + lay["p_len"]] = xg[nu_g + lay["p_glob0"]: nu_g + lay["p_glob0"] + lay["p_len"]]

        def halo(base, pd, o0, o1, gb, ga):
            reqs = []
            t = torch.from_numpy(loc)
            if rank > 0:
                reqs.append(dist.irecv(t[base + (o0 - gb) * pd: base + o0 * pd], src=rank - 1))
                reqs.append(dist.isend(t[base + o0 * pd: base + (o0 + ga) * pd].clone(), dst=rank - 1))
            if rank < world - 1:
                reqs.append(dist.irecv(t[base + o1 * pd: base + (o1 + ga) * pd], src=rank + 1))
                reqs.append(dist.isend(t[base + (o1 - gb) * pd: base + o1 * pd].clone(), dst=rank + 1))
            for r in reqs:
                r.wait()
        last = rank == world - 1
        ou0, ou1 = 2 * (lay["k0"] - lay["e0"]), 2 * (lay["k1"] - lay["e0"]) + (1 if last else 0)
        op0, op1 = lay["k0"] - lay["e0"], lay["k1"] - lay["e0"] + (1 if last else 0)
        halo(0, pu, ou0, ou1, 2, 1)
        halo(nu_loc, pp, op0, op1, 1, 1)
        # every plane an owned row reads must now hold the global values: u planes [2k0-2, 2k1], p planes [k0-1, k1]
        zu0, zu1 = max(0, 2 * lay["k0"] - 2), min(2 * mz, 2 * lay["k1"])
        for z in range(zu0, zu1 + 1):
            zl = z - 2 * lay["e0"]
            assert np.array_equal(loc[zl * pu:(zl + 1) * pu], xg[z * pu:(z + 1) * pu]), ("u plane", z)
        zp0, zp1 = max(0, lay["k0"] - 1), min(mz, lay["k1"])
        for z in range(zp0, zp1 + 1):
            zl = z - lay["e0"]
            assert np.array_equal(loc[nu_loc + zl * pp: nu_loc + (zl + 1) * pp], xg[nu_g + z * pp: nu_g + (z + 1) * pp]), ("p plane", z)
        # 4. replay the plane-distributed coarse levels (Level::pdist): ownership from xsb_pdist_range, products exchange their
        #    OUTPUT planes (one below, one above), restriction to the next level needs nothing further, and the gather into a
        #    replicated level is one broadcast per rank of the planes it restricted
        for depth in range(0, 3):
            rng = [X.pdist_range(mz, world, r, depth) for r in range(world)]
            if mz % (1 << depth) or min(b - a for a, b in rng) < 1:
                break
            nzl = mz // (1 << depth) + 1
            assert rng[0][0] == 0 and rng[-1][1] == nzl and all(rng[r][1] == rng[r + 1][0] for r in range(world - 1)), ("ranges tile the level", depth, rng)
            p0, p1 = rng[rank]; pd = 5
            want = np.sin(0.13 * np.arange(nzl * pd) + depth)
            v = np.full(nzl * pd, np.nan); v[p0 * pd:p1 * pd] = want[p0 * pd:p1 * pd]          # "product": owned rows only
            ops = []; ghosts = []
            if rank > 0:
                ops.append(dist.P2POp(dist.isend, torch.from_numpy(v[p0 * pd:(p0 + 1) * pd].copy()), rank - 1)); gb = torch.empty(pd, dtype=torch.float64); ops.append(dist.P2POp(dist.irecv, gb, rank - 1)); ghosts.append((p0 - 1, gb))
            if rank < world - 1:
                ops.append(dist.P2POp(dist.isend, torch.from_numpy(v[(p1 - 1) * pd:p1 * pd].copy()), rank + 1)); ga = torch.empty(pd, dtype=torch.float64); ops.append(dist.P2POp(dist.irecv, ga, rank + 1)); ghosts.append((p1, ga))
            for w_ in dist.batch_isend_irecv(ops):
                w_.wait()
            for z, t in ghosts:
                v[z * pd:(z + 1) * pd] = t.numpy()
            lo, hi = max(p0 - 1, 0), min(p1 + 1, nzl)
            assert np.array_equal(v[lo * pd:hi * pd], want[lo * pd:hi * pd]), ("planes current after the exchange", depth)
            # restriction stencil of the owned coarse planes stays inside the current planes
            K0, K1 = (p0 + 1) // 2, (p1 + 1) // 2
            assert (K0, K1) == X.pdist_range(mz, world, rank, depth + 1)
            for K in range(K0, K1):
                for z in range(max(2 * K - 1, 0), min(2 * K + 1, nzl - 1) + 1):
                    assert lo <= z < hi, ("restriction reads current planes", depth, K, z)
            # prolongation of the owned planes and of the two neighbour planes reads coarse planes [K0-1, K1]
            for z in range(lo, hi):
                for cz in ((z // 2,) if z % 2 == 0 else (z // 2, z // 2 + 1)):
                    assert K0 - 1 <= cz <= K1, ("prolongation reads current coarse planes", depth, z)
        dist.barrier(); dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:   # report instead of hanging the parent
        q.put((rank, repr(e)))


@pytest.mark.parametrize("world,mesh", [(2, (4, 3, 8)), (2, (2, 2, 5)), (4, (3, 3, 8)), (3, (2, 2, 7))])
def test_slab_layout_and_halo_plane_arithmetic_gloo(world, mesh):
    ctx = mp.get_context("spawn")
    q = ctx.Queue(); port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mesh[0], mesh[1], mesh[2], q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


def test_slab_layout_single_rank_is_the_whole_mesh():
    lay = X.slab_layout(3, 4, 4, 4, 1, 0)
    assert (lay["k0"], lay["k1"], lay["e0"], lay["e1"], lay["u_off"]) == (0, 4, 0, 4, 0)
    assert lay["u_len"] == 3 * 9 * 9 * 9 and lay["p_len"] == 125 and lay["p_off"] == lay["u_len"]
