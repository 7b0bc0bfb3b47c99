"""CPU-only tests of the product's host side: the C-ABI library loads and exports every declared symbol,
the integer index maps are bit-exact against the oracle, option handling and error behaviour mirror the
reference (no compute calls: there is no GPU here and the library has no CPU path)."""
import os
import re

import numpy as np
import pytest

import exsaddle_b200 as X
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "exsaddle_b200.h")).read()
    declared = set(re.findall(r"\b(xsb_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("xsb_ctx")
    assert len(declared) >= 40
    L = X.lib()
    for name in sorted(declared):
        assert hasattr(L, name), "symbol %s declared in include/exsaddle_b200.h but not exported" % name


def test_no_torch_types_in_abi():
    hdr = open(os.path.join(ROOT, "include", "exsaddle_b200.h")).read()
    assert "torch" not in hdr.replace("no torch", "") and "at::" not in hdr and "std::" not in hdr


@pytest.mark.parametrize("nsd,m", [(2, (4, 4, 1)), (2, (3, 5, 1)), (3, (2, 2, 2)), (3, (4, 7, 5)), (3, (1, 1, 1)), (3, (3, 2, 4))])
def test_pattern_rows_bit_exact_vs_oracle(nsd, m):
    p = O.Problem("-mx %d -my %d -mz %d -model %d" % (m[0], m[1], m[2], 0), nsd=nsd)
    A = p.A()
    for row in range(p.n):
        cols = X.pattern_row(nsd, m[0], m[1], m[2], row)
        assert np.array_equal(cols, A.ja[A.ia[row]:A.ia[row + 1]]), row
    assert X.prealloc_total(nsd, *m) == p.prealloc


@pytest.mark.parametrize("nsd,lame,model,fs", [(2, 0, 0, 0), (2, 0, 0, 1), (3, 0, 0, 0), (3, 0, 0, 1), (3, 0, 11, 0), (3, 1, 8, 0),
                                              (3, 1, 9, 0), (3, 1, 12, 0), (2, 1, 9, 0), (3, 1, 6, 0), (2, 0, 101, 0)])
def test_bc_list_matches_oracle(nsd, lame, model, fs):
    m = (4, 4, 4)
    p = O.Problem("-mx 4 -model %d %s" % (model, "-freesliphack" if fs else ""), nsd=nsd, lame=bool(lame))
    idx, val = X.bc_list(nsd, lame, model, fs, *m)
    oi, ov = p.bc()
    assert np.array_equal(idx, oi)
    if model != 101:   # MMS values are filled from coordinates during assembly
        assert np.array_equal(val, ov)


def test_compression_quirk_noncubic():
    """models.c:270 tests si+ni against the y count: on a non-square mesh the x-max face is not constrained."""
    idx, val = X.bc_list(3, 1, 9, 0, 4, 3, 3)
    p = O.Problem("-mx 4 -my 3 -mz 3 -model 9", nsd=3, lame=True)
    assert np.array_equal(idx, p.bc()[0]) and len(idx) == 3 * 7 * 7


def test_mg_level_dims_and_errors():
    assert X.mg_level_dims(3, 6, 6, 6, 3, 2) == (13, 13, 13)
    assert X.mg_level_dims(3, 6, 6, 6, 3, 1) == (7, 7, 7)
    assert X.mg_level_dims(3, 6, 6, 6, 3, 0) == (4, 4, 4)
    assert X.mg_level_dims(2, 32, 32, 1, 3, 0) == (17, 17, 1)
    assert X.mg_level_dims(3, 64, 64, 64, 6, 0) == (5, 5, 5)
    with pytest.raises(X.XsbError):
        X.mg_level_dims(3, 6, 6, 6, 4, 0)     # 4 nodes cannot be coarsened again


def test_slab_ranges_partition_the_mesh():
    for mz, P in [(128, 8), (64, 4), (10, 3), (7, 7)]:
        rs = [X.slab_range(mz, P, r) for r in range(P)]
        assert rs[0][0] == 0 and rs[-1][1] == mz
        assert all(rs[i][1] == rs[i + 1][0] for i in range(P - 1))
        sizes = [b - a for a, b in rs]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(X.XsbError):
        X.slab_range(3, 4, 0)


def test_options_and_banner_without_device(kat):
    c = kat["exSaddle3d_ar_1"]
    s = X.ExSaddle(nsd=3)
    s.set_options(c["options"].replace("-options_file abf.opts", " ".join(kat["_abf_opts"])))
    assert s.banner().rstrip("\n").split("\n") == c["banner"]
    s2 = X.ExSaddle("-model 6 -mx 4", nsd=3, lame=True)
    assert s2.banner().rstrip("\n").split("\n") == kat["exSaddle3d_lame_1"]["banner"]
    s3 = X.ExSaddle("-model 11 -size_x 0.1 -mx 6", nsd=3)
    assert s3.banner().rstrip("\n").split("\n") == kat["exSaddle3d_pseudoice_1"]["banner"]


def test_unsupported_models_error_like_reference():
    with pytest.raises(X.XsbError) as e:
        X.ExSaddle("-model 3", nsd=3).banner()
    assert "not implemented" in str(e.value)          # models.c:1521
    with pytest.raises(X.XsbError) as e:
        X.ExSaddle("-model 2 -sinker_n 9", nsd=3).banner()
    assert "Too many sinkers" in str(e.value)         # models.c:1041
    with pytest.raises(X.XsbError) as e:
        X.ExSaddle("-model 2 -sinker_r 0.06", nsd=3).banner()
    assert "Sinker Radius too big" in str(e.value)    # models.c:1044


@pytest.mark.skipif(X.device_available(), reason="checks the no-GPU behaviour")
def test_compute_calls_fail_loudly_without_gpu():
    s = X.ExSaddle("-mx 2 -model 0", nsd=3)
    with pytest.raises(X.XsbError) as e:
        s.assemble()
    assert e.value.code == -6 and "no CPU path" in str(e.value)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under exsaddle_b200/ or include/ may reference it."""
    for d in ("exsaddle_b200", "include"):
        for root, _, files in os.walk(os.path.join(ROOT, d)):
            if "build" in root.split(os.sep) or "__pycache__" in root:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp")):
                    txt = open(os.path.join(root, f)).read()
                    assert "oracle" not in txt.lower().replace("no oracle", ""), os.path.join(root, f)
                    assert "libxo" not in txt and "xo_" not in txt, os.path.join(root, f)


# ------------------------------------------------------------------ bench.py host logic (no GPU)
def test_bench_reference_arm_line_shape():
    """`bench.py --impl reference` (the CPU arm the driver runs first): one JSON line with the contract's keys, on a tiny mesh."""
    import json, subprocess, sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--mx", "8", "--levels", "3", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.split("\n") if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "higher_is_better", "scaling", "dtype", "config", "cpu_baseline", "e2e"):
        assert k in d
    assert d["impl"] == "reference" and d["higher_is_better"] is False and d["dtype"] == "f64" and d["unit"] == "s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "outer FGMRES" in d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_bench_algorithmic_bytes_and_path_choice():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py")); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    # BAIJ(3) A00 at 64^3 = 10.372 GB for a plain product (x, y streams), CSR full A = 17.183 GB
    nun = 129 ** 3
    per_dir = sum(3 if (i % 2 or i in (0, 128)) else 5 for i in range(129))   # coupling width of node i along one direction (range_uu)
    nblk = per_dir ** 3
    bytes_plain = b.a00_bytes((3 * nun, 3 * nun, 9 * nblk, 3), [1, 0, 0, 0])
    assert abs(bytes_plain - 10.372e9) < 0.01e9
    assert b.a00_bytes((3 * nun, 3 * nun, 9 * nblk, 3), [0, 0, 0, 1]) > bytes_plain     # Chebyshev step streams 3 more vectors
    assert abs(b.a00_csr_bytes(64) - 14.709e9) < 0.001e9 and abs(b.a00_csr_bytes(32) - 1.850e9) < 0.001e9   # A00 alone, AIJ layout
    # 32-bit PetscInt: the assembled operator needs nnz(A) per rank < 2^31 (5420 nnz per element at large m)
    assert 5420.0 * 64 ** 3 < 2.0e9 and 5420.0 * 128 ** 3 / 4 > 2.0e9 and 5420.0 * 128 ** 3 / 8 < 2.0e9


# ------------------------------------------------------------------ PETSc binary writers (no GPU)
def test_petsc_binary_writers_byte_layout(tmp_path):
    """The format PetscViewerBinaryOpen + MatView / VecView write (exSaddle_io.c:61-88): big-endian, class id first."""
    import struct
    ia = np.array([0, 2, 3, 5], np.int32); ja = np.array([0, 2, 1, 0, 2], np.int32); a = np.array([1.5, -2.0, 3.25, 4.0, 1e-300])
    pm, pv = str(tmp_path / "operator_0.petscbin"), str(tmp_path / "solution.petscbin")
    X.write_petsc_mat(pm, ia, ja, a, (3, 3))
    want = struct.pack(">4i", 1211216, 3, 3, 5) + struct.pack(">3i", 2, 1, 2) + struct.pack(">5i", *ja) + struct.pack(">5d", *a)
    assert open(pm, "rb").read() == want
    x = np.array([0.0, -1.0, 2.5e10, np.pi])
    X.write_petsc_vec(pv, x)
    assert open(pv, "rb").read() == struct.pack(">2i", 1211214, 4) + struct.pack(">4d", *x)
    kind, (ia2, ja2, a2, shape) = X.read_petsc_binary(pm)
    assert kind == "Mat" and shape == (3, 3) and np.array_equal(ia2, ia) and np.array_equal(ja2, ja) and np.array_equal(a2, a)
    kind, x2 = X.read_petsc_binary(pv)
    assert kind == "Vec" and np.array_equal(x2, x)


def test_petsc_binary_round_trip_of_the_oracle_operator(tmp_path):
    from oracle import oracle as O
    o = O.Problem("-model 6 -mx 3 -my 2 -mz 2 -eta1 10", nsd=3)
    A = o.A()
    path = str(tmp_path / "operator_0.petscbin")
    X.write_petsc_mat(path, A.ia, A.ja, A.a, A.shape)
    kind, (ia, ja, a, shape) = X.read_petsc_binary(path)
    assert kind == "Mat" and shape == A.shape and np.array_equal(ia, A.ia) and np.array_equal(ja, A.ja) and np.array_equal(a, A.a)
    assert os.path.getsize(path) == 16 + 4 * o.n + 12 * o.nnz


def test_vts_writer_round_trip(tmp_path):
    """ViewFields container (exSaddle_io.c:128-177): the interleaved velocity vector as scalar point fields u, v, w on the node
    lattice with uniform coordinates; read back with an independent parser of the VTK XML appended-raw layout."""
    nx, ny, nz, nsd = 5, 3, 4, 3
    n = nx * ny * nz
    xu = np.arange(nsd * n, dtype=np.float64) * 0.5 - 7.0        # node-major, component fastest (DMDA dof ordering)
    path = str(tmp_path / "uvw.vts")
    X.write_vts(path, (nx, ny, nz), (0.25, 0.5, 0.125), ["u", "v", "w"], xu, 1, nsd)
    dims, pts, fields = X.read_vts(path)
    assert dims == (nx, ny, nz) and sorted(fields) == ["u", "v", "w"]
    for c, name in enumerate(("u", "v", "w")):
        assert np.array_equal(fields[name], xu[c::nsd])
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    node = (i + nx * (j + ny * k)).ravel()
    assert np.array_equal(pts[node], np.stack([0.25 * i.ravel(), 0.5 * j.ravel(), 0.125 * k.ravel()], axis=1))
    head = open(path, "rb").read(400).decode(errors="replace")
    assert 'type="StructuredGrid"' in head and 'byte_order="LittleEndian"' in head and 'WholeExtent="0 4 0 2 0 3"' in head


def test_bench_reference_arm_other_ranks_exit_without_work():
    """Under torchrun (N > 1) rank 0 alone runs the CPU arm; the other ranks exit 0 and print nothing."""
    import subprocess, sys
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], cwd=ROOT, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.parametrize("nsd,lame,opts", [(3, False, "-model 6 -mx 4 -my 3 -mz 2 -eta1 100"), (2, False, "-model 0 -mx 5 -my 3 -size_x 2.0"), (3, True, "-model 6 -mx 3 -mu1 10 -lambda1 5")])
def test_gradient_block_from_the_products_1d_tables_equals_the_oracles_assembled_block(nsd, lame, opts):
    """The tables behind the matrix-free gradient / divergence products (host code of the library, no GPU): A01 rebuilt from them as a
    sum of Kronecker products equals the oracle's assembled A[u,p] to 1e-14 (constrained rows removed on both sides), and the
    pressure-side tables are the transpose view of the velocity-side ones."""
    import scipy.sparse as sp
    from oracle import oracle as O
    p = O.Problem(opts, nsd=nsd, lame=lame); o = O.parse_options(opts)
    mx = int(o.get("mx", 4)); mesh = [mx, int(o.get("my", mx)), int(o.get("mz", mx))][:nsd]
    size = [float(o.get("size_" + "xyz"[d], 1.0)) for d in range(nsd)]
    nu = p.nu; A01 = p.A().scipy().tocsr()[:nu, nu:]
    fac = []
    for d in range(nsd):
        uP, uM, uG, pM, pG = X.grad_line_tables(mesh[d], size[d] / (2 * mesh[d]))
        N, P = 2 * mesh[d] + 1, mesh[d] + 1
        M = np.zeros((N, P)); G = np.zeros((N, P))
        for i in range(N):
            for a in range(3):
                if uP[i, a] >= 0:
                    M[i, uP[i, a]] += uM[i, a]; G[i, uP[i, a]] += uG[i, a]
        for Pn in range(P):            # the divergence kernel's view of the same numbers
            for di in range(5):
                i = 2 * Pn - 2 + di
                assert (pM[Pn, di], pG[Pn, di]) == ((M[i, Pn], G[i, Pn]) if 0 <= i < N else (0.0, 0.0))
        fac.append((sp.csr_matrix(M), sp.csr_matrix(G)))
    blocks = []
    for c in range(nsd):               # component c: derivative factor in its own direction; node index = i + NX (j + NY k): kron(z, y, x)
        K = None
        for d in reversed(range(nsd)):
            f = fac[d][1] if d == c else fac[d][0]
            K = f if K is None else sp.kron(K, f, format="csr")
        blocks.append(-K)
    n_nodes = blocks[0].shape[0]
    B = sp.lil_matrix((nu, A01.shape[1]))
    rows = np.arange(n_nodes) * nsd
    B = sp.vstack([blk for blk in blocks]).tocsr()                      # rows grouped by component ...
    perm = np.concatenate([np.arange(n_nodes) + c * n_nodes for c in range(nsd)]).reshape(nsd, n_nodes).T.ravel()
    B = B[perm]                                                           # ... back to node-major, component-fastest order
    bi, _ = p.bc(); keep = np.ones(nu, bool); keep[bi[bi < nu]] = False
    D = (sp.diags(keep.astype(float)) @ B - A01).tocoo()
    assert np.max(np.abs(D.data), initial=0.0) <= 1e-14 * np.max(np.abs(A01.data))


@pytest.mark.parametrize("nranks", [2, 3])
def test_gradient_tables_on_a_slab_lattice_give_the_global_rows_of_the_owned_planes(nranks):
    """Slabs apply the gradient / divergence stencils on the rank's LOCAL lattice (element layers [k0-2, k1+1)), whose ends are treated
    like mesh boundaries.  Claim: for the planes a rank OWNS this gives the rows of the global blocks -- every element around an owned
    node lies inside the local lattice.  Checked with the product's tables (z direction: local element count) against the oracle's
    global A[u,p] and A[p,u], rank by rank."""
    import scipy.sparse as sp
    from oracle import oracle as O
    mx, my, mz = 3, 2, 6
    p = O.Problem("-model 6 -mx %d -my %d -mz %d -eta1 10" % (mx, my, mz), nsd=3)
    nu = p.nu; A = p.A().scipy().tocsr(); A01 = A[:nu, nu:].toarray(); A10 = A[nu:, :nu].toarray()
    NX, NY, PX, PY = 2 * mx + 1, 2 * my + 1, mx + 1, my + 1
    bi, _ = p.bc(); free = np.ones(nu, bool); free[bi[bi < nu]] = False

    def dense(m, h):
        uP, uM, uG, _, _ = X.grad_line_tables(m, h)
        M = np.zeros((2 * m + 1, m + 1)); G = np.zeros_like(M)
        for i in range(2 * m + 1):
            for a in range(3):
                if uP[i, a] >= 0:
                    M[i, uP[i, a]] += uM[i, a]; G[i, uP[i, a]] += uG[i, a]
        return M, G
    Mx, Gx = dense(mx, 0.5 / mx); My, Gy = dense(my, 0.5 / my)
    for r in range(nranks):
        lay = X.slab_layout(3, mx, my, mz, nranks, r)
        e0, e1, k0, k1 = lay["e0"], lay["e1"], lay["k0"], lay["k1"]
        last = r == nranks - 1
        Mz, Gz = dense(e1 - e0, 0.5 / mz)                                   # the local lattice in z
        for kl in range(2 * (k0 - e0), 2 * (k1 - e0) + (1 if last else 0)):  # owned velocity planes (local index)
            kg = kl + 2 * e0
            for j in range(NY):
                for i in range(NX):
                    node = i + NX * (j + NY * kg)
                    for c in range(3):
                        fx, fy, fz = (Gx if c == 0 else Mx)[i], (Gy if c == 1 else My)[j], (Gz if c == 2 else Mz)[kl]
                        row = np.zeros((mz + 1, PY, PX))
                        row[e0:e1 + 1] = -np.einsum("k,j,i->kji", fz, fy, fx)
                        want = A01[3 * node + c].reshape(mz + 1, PY, PX)
                        got = row if free[3 * node + c] else np.zeros_like(row)
                        assert np.max(np.abs(got - want)) <= 1e-14, (r, kl, j, i, c)
        for Pl in range(k0 - e0, k1 - e0 + (1 if last else 0)):            # owned pressure planes: rows of A10 = columns of the same stencil
            Pg = Pl + e0
            for Pj in range(PY):
                for Pi in range(PX):
                    want = A10[Pi + PX * (Pj + PY * Pg)].reshape(2 * mz + 1, NY, NX, 3)
                    got = np.zeros_like(want)
                    for c in range(3):
                        fx, fy, fz = (Gx if c == 0 else Mx)[:, Pi], (Gy if c == 1 else My)[:, Pj], (Gz if c == 2 else Mz)[:, Pl]
                        got[2 * e0:2 * e1 + 1, :, :, c] = -np.einsum("k,j,i->kji", fz, fy, fx)
                    got = got * free.reshape(2 * mz + 1, NY, NX, 3)
                    assert np.max(np.abs(got - want)) <= 1e-14, (r, Pl, Pj, Pi)
