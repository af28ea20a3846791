#!/usr/bin/env python
"""Extract the reference's own gold results for the hot path into small fixtures.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):  python tests/golden/make_golden.py
Writes tests/golden/*.npz.  Only NUMBERS from the reference's gold files (test data) are
stored - no reference source code.

Sources (relative to the reference root):
  test/tests/cahnhilliard/gold/cahnhilliard_out.e   Exodus/NetCDF-3; nodal `c`, elemental `mu`
  test/tests/solvers/gold/{diagonal,coupled,nl_coupled}_*.csv   postprocessor CSVs
  test/tests/solvers/gold/etdrk4_diffusion_rmse.csv
  test/tests/mechanics/gold/{mech3d,mech}.h5        HDF5, one deflate chunk per dataset
  test/tests/gradient/gold/*.csv, test/tests/tensor_compute/gold/backandforth_out.csv
"""
import os
import zlib

import numpy as np
from scipy.io import netcdf_file

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def exodus_ch2d():
    f = netcdf_file(f"{REF}/test/tests/cahnhilliard/gold/cahnhilliard_out.e", "r", mmap=False)
    v = f.variables
    n, L = 20, 3.0
    dx = L / n
    x, y = v["coordx"][:].copy(), v["coordy"][:].copy()
    nod = v["vals_nod_var1"][:].copy()           # [time, node]
    elv = v["vals_elem_var1eb1"][:].copy()       # [time, elem]
    conn = v["connect1"][:].copy() - 1
    times = v["time_whole"][:].copy()
    nt = nod.shape[0]
    # ProjectTensorAux (src/auxkernels/ProjectTensorAux.C:38-60): nodal value at node p is
    # buffer[int((p+dx/2)/dx) % n]; elemental value at centroid c is buffer[int(c/dx) % n]
    c = np.full((nt, n, n), np.nan)
    ii = np.floor((x + dx / 2) / dx + 1e-9).astype(int) % n
    jj = np.floor((y + dx / 2) / dx + 1e-9).astype(int) % n
    for t in range(nt):
        # wrap-around nodes carry duplicate values; interior assignment wins consistently
        c[t, ii, jj] = nod[t]
    cx, cy = x[conn].mean(1), y[conn].mean(1)
    ei = np.floor(cx / dx).astype(int) % n
    ej = np.floor(cy / dx).astype(int) % n
    mu = np.full((nt, n, n), np.nan)
    for t in range(nt):
        mu[t, ei, ej] = elv[t]
    assert not np.isnan(c).any() and not np.isnan(mu).any()
    np.savez_compressed(f"{OUT}/ch2d_exodus.npz", c=c, mu=mu, time=times)
    print("ch2d_exodus", c.shape, mu.shape)


def exodus_fields(path, n, L, frames=None):
    """Nodal variable 1 -> c, elemental variable 1 -> mu of a ProjectTensorAux exodus file on an
    n^dim grid of extent L (2-D or 3-D), mapped back to the tensor layout."""
    f = netcdf_file(path, "r", mmap=False)
    v = f.variables
    dim = 3 if "coordz" in v and np.abs(v["coordz"][:]).max() > 0 else 2
    dx = L / n
    xyz = [v[k][:].copy() for k in ("coordx", "coordy", "coordz")[:dim]]
    nod, elv = v["vals_nod_var1"][:].copy(), v["vals_elem_var1eb1"][:].copy()
    conn = v["connect1"][:].copy() - 1
    times = v["time_whole"][:].copy()
    sel = list(range(nod.shape[0])) if frames is None else frames
    ni = tuple(np.floor((a + dx / 2) / dx + 1e-9).astype(int) % n for a in xyz)
    ei = tuple(np.floor(a[conn].mean(1) / dx).astype(int) % n for a in xyz)
    c = np.full((len(sel),) + (n,) * dim, np.nan)
    mu = np.full((len(sel),) + (n,) * dim, np.nan)
    for k, t in enumerate(sel):
        c[(k,) + ni] = nod[t]
        mu[(k,) + ei] = elv[t]
    assert not np.isnan(c).any() and not np.isnan(mu).any()
    return c, mu, times[sel]


def exodus_more():
    c, mu, t = exodus_fields(f"{REF}/test/tests/cahnhilliard/gold/map_to_aux_3d.e", 5, 3.0)
    np.savez_compressed(f"{OUT}/ch3d_map_to_aux_exodus.npz", c=c, mu=mu, time=t)
    print("ch3d_map_to_aux_exodus", c.shape)
    frames = [0, 1, 2, 5, 10, 20, 50, 100]
    c, mu, t = exodus_fields(f"{REF}/test/tests/cahnhilliard/gold/cahnhilliard_explicit_out.e", 50, 3.0, frames)
    np.savez_compressed(f"{OUT}/ch2d_explicit_exodus.npz", c=c, mu=mu, time=t, frames=np.array(frames))
    print("ch2d_explicit_exodus", c.shape, t)
    frames = [0, 1, 5, 20]
    for name in ("sharp", "houli"):
        c, mu, t = exodus_fields(f"{REF}/test/tests/cahnhilliard/gold/{name}.e", 50, 3.0, frames)
        np.savez_compressed(f"{OUT}/ch2d_explicit_{name}_exodus.npz", c=c, mu=mu, time=t, frames=np.array(frames))
        print(name, c.shape, t)


def read_csv(path):
    with open(path) as fh:
        header = fh.readline().strip().split(",")
        rows = [[float(x) for x in line.strip().split(",")] for line in fh if line.strip()]
    return header, np.array(rows)


def solver_csvs():
    out = {}
    gd = f"{REF}/test/tests/solvers/gold"
    for fn in sorted(os.listdir(gd)):
        if fn.startswith(("diagonal_", "etdrk4", "coupled_", "nl_coupled_")):
            h, a = read_csv(f"{gd}/{fn}")
            out[fn[:-4]] = a
            out[fn[:-4] + "__header"] = np.array(h)
    for fn, key in [("test/tests/gradient/gold/gradient_out.csv", "gradient_out"),
                    ("test/tests/gradient/gold/gradient_square_out.csv", "gradient_square_out"),
                    ("test/tests/tensor_compute/gold/backandforth_out.csv", "backandforth_out"),
                    ("test/tests/parsed_tensor/gold/local_vars_derivative_out.csv",
                     "local_vars_derivative_out"),
                    ("test/tests/postprocessors/gold/interface_velocity_out.csv", "interface_velocity_out"),
                    ("test/tests/histogram/gold/test_out_hist_0001.csv", "histogram_out_hist_0001")]:
        if os.path.exists(f"{REF}/{fn}"):
            h, a = read_csv(f"{REF}/{fn}")
            out[key] = a
            out[key + "__header"] = np.array(h)
    np.savez_compressed(f"{OUT}/csv_golds.npz", **out)
    print("csv_golds", [k for k in out if not k.endswith("__header")])


def zlib_streams(path):
    """Every dataset is one deflate-9 chunk (XDMFTensorOutput.C:597-622): scan for zlib
    headers and inflate."""
    data = open(path, "rb").read()
    pos, found = 0, []
    while True:
        i = data.find(b"\x78\xda", pos)
        if i < 0:
            break
        try:
            d = zlib.decompressobj()
            out = d.decompress(data[i:])
            used = len(data) - i - len(d.unused_data)
            if len(out) >= 64 and d.eof:
                found.append(out)
                pos = i + used
                continue
        except zlib.error:
            pass
        pos = i + 1
    return found


def mech3d():
    n = 16
    streams = zlib_streams(f"{REF}/test/tests/mechanics/gold/mech3d.h5")
    # per frame, std::map key order: F_0..F_8, disp_x, disp_y, disp_z, phase, sV
    per = 14
    assert len(streams) % per == 0, len(streams)
    frames = len(streams) // per
    F = np.zeros((frames, n, n, n, 3, 3))
    sV = np.zeros((frames, n, n, n))
    for fr in range(frames):
        blk = streams[fr * per:(fr + 1) * per]
        for k in range(9):
            a = np.frombuffer(blk[k], dtype="<f8").reshape(n, n, n)  # stored [z,y,x]
            F[fr, :, :, :, k // 3, k % 3] = a.transpose(2, 1, 0)
        sV[fr] = np.frombuffer(blk[13], dtype="<f8").reshape(n, n, n).transpose(2, 1, 0)
    # disp_x, disp_y, disp_z: OVERSIZED_NODAL (n+1)^3, stored transposed like the cell data
    disp = np.zeros((frames, n + 1, n + 1, n + 1, 3))
    for fr in range(frames):
        for k in range(3):
            disp[fr, ..., k] = np.frombuffer(streams[fr * per + 9 + k], dtype="<f8").reshape(n + 1, n + 1, n + 1).transpose(2, 1, 0)
    np.savez_compressed(f"{OUT}/mech3d_h5.npz", F=F, sV=sV, disp=disp)
    print("mech3d_h5", F.shape)


def mech2d():
    """test/tests/mechanics/gold/mech.h5 (32^2, 2x2 tensors): per frame F_0..F_3, disp_x, disp_y,
    phase (nodal 33^2), sV."""
    n = 32
    streams = zlib_streams(f"{REF}/test/tests/mechanics/gold/mech.h5")
    per = 8
    assert len(streams) % per == 0, len(streams)
    frames = len(streams) // per
    F = np.zeros((frames, n, n, 2, 2))
    sV = np.zeros((frames, n, n))
    for fr in range(frames):
        blk = streams[fr * per:(fr + 1) * per]
        for k in range(4):
            a = np.frombuffer(blk[k], dtype="<f8").reshape(n, n)  # stored [y,x]
            F[fr, :, :, k // 2, k % 2] = a.T
        sV[fr] = np.frombuffer(blk[7], dtype="<f8").reshape(n, n).T
    disp = np.zeros((frames, n + 1, n + 1, 2))
    for fr in range(frames):
        for k in range(2):
            disp[fr, ..., k] = np.frombuffer(streams[fr * per + 4 + k], dtype="<f8").reshape(n + 1, n + 1).T
    np.savez_compressed(f"{OUT}/mech2d_h5.npz", F=F, sV=sV, disp=disp)
    print("mech2d_h5", F.shape)


def rotating_grain():
    """test/tests/tensor_compute/gold/rotating_grain_secant.h5: psi (40^2, CELL, transpose = false) at
    the initial condition and after each of the 10 steps."""
    streams = zlib_streams(f"{REF}/test/tests/tensor_compute/gold/rotating_grain_secant.h5")
    psi = np.stack([np.frombuffer(s, dtype="<f8").reshape(40, 40) for s in streams])
    np.savez_compressed(f"{OUT}/rotating_grain_secant_h5.npz", psi=psi)
    print("rotating_grain_secant_h5", psi.shape)


def ch2d_slab_rank1():
    """test/tests/cahnhilliard/gold/cahnhilliard.rank0001.h5 (xdmf_output_hdf5_parallel: 2 ranks, FFT_SLAB, CELL, transpose =
    false): rank 1's local part of c, [20][10] (real space is split along y), at the initial condition and after
    each of the 10 steps, in time order."""
    streams = zlib_streams(f"{REF}/test/tests/cahnhilliard/gold/cahnhilliard.rank0001.h5")
    assert len(streams) == 11
    c = np.stack([np.frombuffer(s, dtype="<f8").reshape(20, 10) for s in streams])
    np.savez_compressed(f"{OUT}/ch2d_slab_rank1_h5.npz", c=c)
    print("ch2d_slab_rank1_h5", c.shape)


def ch2d_xdmf_h5():
    """test/tests/cahnhilliard/gold/cahnhilliard.h5 (xdmf_output_hdf5: c as NODE data [21][21], mu as CELL data [20][20],
    transpose = false, per frame c then mu): frames 0, 1 and 10."""
    streams = zlib_streams(f"{REF}/test/tests/cahnhilliard/gold/cahnhilliard.h5")
    assert [len(s) // 8 for s in streams] == [441, 400] * 11
    keep = [0, 1, 10]
    c = np.stack([np.frombuffer(streams[2 * k], dtype="<f8").reshape(21, 21) for k in keep])
    mu = np.stack([np.frombuffer(streams[2 * k + 1], dtype="<f8").reshape(20, 20) for k in keep])
    np.savez_compressed(f"{OUT}/ch2d_xdmf_h5.npz", c_node=c, mu_cell=mu, frames=np.array(keep))
    print("ch2d_xdmf_h5", c.shape, mu.shape)


def smooth_rectangle():
    """test/tests/tensor_compute/gold/smooth_rectangle.h5: rectangle_cos, rectangle_sharp, rectangle_tanh (100^2,
    datasets in HDF5 name order)."""
    streams = zlib_streams(f"{REF}/test/tests/tensor_compute/gold/smooth_rectangle.h5")
    assert len(streams) == 3
    cos, sharp, tanh = [np.frombuffer(s, dtype="<f8").reshape(100, 100) for s in streams]
    np.savez_compressed(f"{OUT}/smooth_rectangle_h5.npz", cos=cos, sharp=sharp, tanh=tanh)
    print("smooth_rectangle_h5", cos.shape)


def kks_no_flux():
    """test/tests/kks/gold/KKS_no_flux_bc.h5 (20^2, transpose = false; per frame c, eta, mu, psi in
    std::map key order) and KKS_no_flux_bc_out.csv."""
    streams = zlib_streams(f"{REF}/test/tests/kks/gold/KKS_no_flux_bc.h5")
    a = np.stack([np.frombuffer(s, dtype="<f8").reshape(20, 20) for s in streams])
    # file order of the chunks: frame 0 (c, eta, mu, psi), the ten later copies of the static psi, then
    # (c, eta, mu) of steps 1..10
    assert a.shape[0] == 44 and all(np.array_equal(a[3], a[j]) for j in range(4, 14))
    frames = np.zeros((11, 4, 20, 20))
    frames[0] = a[0:4]
    for k in range(1, 11):
        frames[k, 0:3] = a[14 + 3 * (k - 1):17 + 3 * (k - 1)]
        frames[k, 3] = a[3]
    a = frames
    h, csv = read_csv(f"{REF}/test/tests/kks/gold/KKS_no_flux_bc_out.csv")
    np.savez_compressed(f"{OUT}/kks_no_flux_bc.npz", c=a[:, 0], eta=a[:, 1], mu=a[:, 2], psi=a[:, 3], csv=csv, csv_header=np.array(h))
    print("kks_no_flux_bc", a.shape, csv.shape)


def xmf_gold():
    """test/tests/cahnhilliard/gold/cahnhilliard.xmf (XMLDiff gold of the XDMF output), kept verbatim."""
    import shutil
    shutil.copyfile(f"{REF}/test/tests/cahnhilliard/gold/cahnhilliard.xmf", f"{OUT}/cahnhilliard_gold.xmf")
    print("cahnhilliard_gold.xmf")


if __name__ == "__main__":
    exodus_ch2d()
    solver_csvs()
    mech3d()
    mech2d()
    rotating_grain()
    smooth_rectangle()
    ch2d_xdmf_h5()
    ch2d_slab_rank1()
    kks_no_flux()
    exodus_more()
    xmf_gold()
