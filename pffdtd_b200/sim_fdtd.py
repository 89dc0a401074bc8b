"""Host run loop: the drop-in for the reference's simulation step.

Mirrors the reference's two front ends for this path --

* the Python engine ``python3 -m fdtd.sim_fdtd --data_dir D [--nsteps N]`` (python/fdtd/sim_fdtd.py:39-937):
  same class name, same method sequence (``load_h5_data, setup_mask, allocate_mem, set_coeffs, checks,
  run_all / run_steps, save_outputs, print_last_samples``) and the same ``--ENGINE:`` log lines;
* the C binaries' ``main`` (c_cuda/fdtd_main.c:35-59): load -> scale_input -> run_sim -> rescale_output ->
  write_outputs -> print_last_samples, whose arithmetic (incl. the input scaling) is what the kernels
  reproduce bit for bit.

It reads the same four ``.h5`` files and writes the same ``sim_outs.h5``, but every
time step runs in the sm_100a kernels of libpffdtd_b200.so.  Launched under ``torchrun`` with
WORLD_SIZE > 1 the grid is split into x-slabs, one process per GPU (gpu_engine.h:516-662), and the halo
planes travel by NCCL send/recv inside the library.

    python -m pffdtd_b200.sim_fdtd --data_dir D [--precision {1,2}] [--nsteps N]
"""
from __future__ import annotations

import os
import time
from pathlib import Path

import numpy as np

from . import parallel
from .engine import Engine, comm_unique_id
from .sim_data import SimData


class SimEngine:
    def __init__(self, data_dir, energy_on=False, nthreads=None, precision=2, device=None, scale=True, quiet=False, timing=False,
                 gpu_folder=False, balance=True):
        self.data_dir = Path(data_dir)
        self.balance = bool(balance)  # cost-weighted slab split (False: the reference's equal-plane split, gpu_engine.h:532-543)
        self.energy_on = bool(energy_on)
        self.H_tot = self.E_lost = self.E_in = None
        self.precision = int(precision)
        self.rank, self.world, local = parallel.dist_env()
        self.device = local if device is None else int(device)
        self.scale = scale
        self.quiet = quiet or self.rank != 0
        self.eng = None
        self.sd_full = None
        self.sd = None
        self.u_out = None
        self.t_elapsed = 0.0
        self.gpu_folder = bool(gpu_folder)  # apply sim_setup's save_folder_gpu transforms (rotate, fold FCC, sort) in memory
        self.timing = bool(timing)  # CUDA events around every air launch -> the reference's air / boundary split lines
        self.t_air = None
        del nthreads  # host threads play no role here; accepted for call compatibility

    def print(self, fstring):
        if not self.quiet:
            print(f"--ENGINE: {fstring}", flush=True)

    # ---- the reference's set-up sequence -------------------------------------------------------
    def load_h5_data(self):
        self.print("loading data..")
        if self.gpu_folder:
            # what sim_setup.py:119-125 + rotate_sim_data.py do to the files on disk, done to the datasets in memory: results equal
            # the reference's run on the gpu folder (not bit for bit the run on the un-rotated folder: the sums change order)
            from . import folder_prep, shoebox
            for fn in ("sim_consts.h5", "vox_out.h5", "comms_out.h5", "sim_mats.h5"):
                if not (self.data_dir / fn).exists():
                    raise FileNotFoundError(f"{fn} doesn't exist in {self.data_dir}")
            if self.energy_on:
                raise ValueError("the energy balance runs on the folder as given (no --gpu_folder)")
            sd = shoebox.sim_data_from_files(folder_prep.gpu_folder(folder_prep.load_folder(self.data_dir)), self.precision)
        else:
            sd = SimData.load(self.data_dir, self.precision)
        if self.scale:
            sd.scale_input()  # fdtd_data.h:879-909
        if self.energy_on and sd.fcc_flag == 2:
            raise ValueError("the energy balance needs the fcc_flag=1 folder (as the reference's Python engine does)")
        if self.world > 1 or sd.fcc_flag == 2:
            if not sd.is_sorted():
                self.print("sorting node lists (the reference needs a sort_sim_data'd folder here)")
            sd = sd.sorted()
        self.sd_full = sd
        self.Nx, self.Ny, self.Nz, self.Nt, self.Nr, self.Ns = sd.Nx, sd.Ny, sd.Nz, sd.Nt, sd.Nr, sd.Ns
        self.fcc = sd.fcc_flag > 0
        self.out_reorder = sd.out_reorder
        self.print(f"Nx={sd.Nx} Ny={sd.Ny} Nz={sd.Nz} Nb={sd.Nb} Nbl={sd.Nbl} Nba={sd.Nba} Ns={sd.Ns} Nr={sd.Nr} Nt={sd.Nt} "
                   f"fcc_flag={sd.fcc_flag} precision={'single' if self.precision == 1 else 'double'}")

    def setup_mask(self):
        pass  # the node mask is built on the device at allocate_mem()

    def allocate_mem(self):
        # slabs of about equal cost (air nodes + weighted boundary / lossy / shell nodes per plane) unless the reference's
        # equal-plane split is asked for; the receiver traces are the same bits either way
        self.planes = None
        if self.world > 1 and self.balance:
            self.planes = SimData.slab_planes(self.sd_full.Nx, self.world, cost=self.sd_full.plane_costs())
            self.print(f"slab planes per rank: {self.planes[1]}")
        self.sd = self.sd_full.slab(self.rank, self.world, planes=self.planes)
        self.eng = Engine(self.sd, self.device)
        if self.world > 1:
            self._comm_init()
        if self.energy_on:
            self.eng.energy_enable()

    def set_coeffs(self):
        pass  # derived in SimData.from_arrays exactly as load_sim_data does (fdtd_data.h:186-194, 424-460)

    def checks(self):
        pass  # SimData.from_arrays and pffdtd_create validate the description

    def _comm_init(self):
        uid = parallel.broadcast_bytes(comm_unique_id() if self.rank == 0 else None, src=0)
        self.eng.comm_init(uid, self.rank, self.world)
        # GPUs of one node: the halo planes go straight into the neighbours' grids over peer memory (NCCL stays as the fallback)
        self.p2p = parallel.connect_peers(self.eng, self.rank, self.world)

    # ---- running ----------------------------------------------------------------------------------
    def run_steps(self, nstart, nsteps):
        self.eng.run_steps(nstart, nsteps)

    def run_all(self, nsteps=1):
        self.print("running..")
        Npts = self.Nx * self.Ny * self.Nz
        t0 = time.perf_counter()
        # batches only bound how far the host runs ahead of the device; there is no per-step sync
        batch = max(int(nsteps), 64)
        if self.timing:
            self.eng.set_option("profile_air", 1)
            self.eng.reset_stats()
        for n in range(0, self.Nt, batch):
            self.run_steps(n, min(batch, self.Nt - n))
        self.eng.sync()
        self.t_elapsed = time.perf_counter() - t0
        self.print(f"Run-time loop: {self.t_elapsed:.6f}, {self.Nt * Npts / 1e6 / max(self.t_elapsed, 1e-12):.2f} MVox/s")
        if self.timing:
            self.t_air = self.eng.stat("air_ms") * 1e-3
            self.eng.set_option("profile_air", 0)
        self._collect()
        self.print_timing()

    def print_timing(self):
        """the three closing lines of the reference's C engines (cpu_engine.h:355-357, gpu_engine.h:1251-1253); "Combined
        (total)" is the published MVPS metric (benchmarks/README.md:9).  The air / boundary split needs `timing=True`
        (CUDA events around every air launch of this rank; everything that is not the air kernel counts as boundary loop)."""
        Npts, Nt, t = self.Nx * self.Ny * self.Nz, self.Nt, max(self.t_elapsed, 1e-12)
        if self.t_air is not None:
            t_air, t_bn = max(self.t_air, 1e-12), max(t - self.t_air, 1e-12)
            self.print(f"Air update: {t_air:.6f}s, {Npts * Nt / 1e6 / t_air:.2f} Mvox/s")
            self.print(f"Boundary loop: {t_bn:.6f}s, {self.sd_full.Nb * Nt / 1e6 / t_bn:.2f} Mvox/s")
        self.print(f"Combined (total): {t:.6f}s, {Npts * Nt / 1e6 / t:.2f} Mvox/s")

    def _collect(self):
        u = parallel.gather_rows(self.eng.read_outputs(0, self.Nt))  # rank order == sorted receiver order
        if self.scale:
            u = self.sd_full.rescale_output(u)  # fdtd_data.h:912-925
        self.u_out = u
        if self.energy_on:
            # every rank summed its own planes; the three series add up across ranks
            # (the engine ran on scale_input()'d sources: energies are quadratic in the field, so undoing the scaling takes infac^2
            # -- the reference's Python engine never scales, sim_fdtd.py:587-620)
            k = self.sd_full.infac ** 2 if self.scale else 1.0
            self.H_tot, self.E_lost, self.E_in = (parallel.sum_arrays(a) * k for a in self.eng.read_energy())

    def print_last_energy(self, Np):
        """sim_fdtd.py:662-669: rel_diff(H_tot+E_lost, E_in) of the last Np steps (common/myfuncs.py:164-165)"""
        self.print("ENERGY")
        for n in range(max(self.Nt - Np, 0), self.Nt):
            self.print(f"normalised energy balance:{energy_balance(self.H_tot, self.E_lost, self.E_in)[n]:.16e}")

    def gather_slice(self, ix=None, iy=None, iz=None):
        """2-D cut through the current state (sim_fdtd.py:640-658), for plots between batches of steps; on the
        checkerboard FCC grid the unused nodes are filled with the mean of their four in-plane neighbours.  One slab only."""
        if self.world > 1:
            raise NotImplementedError("gather_slice reads one engine's grid; run on one GPU for plots")
        u1 = self.eng.read_grid(1)
        if ix is not None:
            uslice, i3 = u1[ix, :, :], ix
        elif iy is not None:
            uslice, i3 = u1[:, iy, :], iy
        elif iz is not None:
            uslice, i3 = u1[:, :, iz], iz
        else:
            raise ValueError("give ix, iy or iz")
        uslice = np.array(uslice) * (self.sd_full.infac if self.scale else 1.0)  # the field of the unscaled problem, as the reference plots it
        if self.sd_full.fcc_flag == 1:
            fcc_fill_plot_holes(uslice, i3)
        return uslice

    def save_outputs(self):
        if self.rank == 0:
            self.sd_full.write_outputs(self.data_dir, self.u_out)
        self.print(f"saved outputs in {self.data_dir}")

    def print_last_samples(self, Np):
        self.print("GRID OUTPUTS")
        for i in range(self.Nr):
            self.print(f"out {i}")
            for n in range(max(self.Nt - Np, 0), self.Nt):
                self.print(f"sample {n}: {self.u_out[self.out_reorder[i], n]:.16e}")

    def close(self):
        if self.eng is not None:
            self.eng.close()
            self.eng = None


def fcc_fill_plot_holes(uslice, i3):
    """nb_fcc_fill_plot_holes (sim_fdtd.py:888-894), in place: interior nodes of odd parity (i1 + i2 + i3) become the mean of their
    four neighbours in the slice"""
    N1, N2 = uslice.shape
    i1, i2 = np.meshgrid(np.arange(1, N1 - 1), np.arange(1, N2 - 1), indexing="ij")
    odd = ((i1 + i2 + i3) % 2) == 1
    a, b = i1[odd], i2[odd]
    # the four neighbours of an odd node are even nodes, which the fill never touches: order does not matter
    uslice[a, b] = 0.25 * (uslice[a + 1, b] + uslice[a - 1, b] + uslice[a, b + 1] + uslice[a, b - 1])
    return uslice


def energy_balance(H_tot, E_lost, E_in):
    """rel_diff(H_tot[n] + E_lost[n], E_in[n]) = (x0 - x1) / 2^floor(log2 x0)   (common/myfuncs.py:164-165); 0 where x0 <= 0"""
    x0 = np.asarray(H_tot) + np.asarray(E_lost)[:len(H_tot)]
    x1 = np.asarray(E_in)[:len(H_tot)]
    out = np.zeros_like(x0)
    ok = x0 > 0
    out[ok] = (x0[ok] - x1[ok]) / 2.0 ** np.floor(np.log2(x0[ok]))
    return out


def run_folder(data_dir, precision=2, device=None, nsteps=1, quiet=True, gpu_folder=False):
    """load -> run -> write sim_outs.h5; returns u_out in file (original receiver) order"""
    eng = SimEngine(data_dir, precision=precision, device=device, quiet=quiet, gpu_folder=gpu_folder)
    eng.load_h5_data()
    eng.setup_mask()
    eng.allocate_mem()
    eng.set_coeffs()
    eng.checks()
    eng.run_all(nsteps)
    eng.save_outputs()
    out = eng.sd_full.reorder_output(eng.u_out)
    eng.close()
    return out


def main(argv=None):
    import argparse
    parser = argparse.ArgumentParser(description="B200 FDTD engine: drop-in for `python -m fdtd.sim_fdtd` / fdtd_main_gpu_*.x")
    parser.add_argument("--data_dir", type=str, help="run directory")
    parser.add_argument("--nsteps", type=int, default=1, help="run in batches of steps")
    parser.add_argument("--nthreads", type=int, default=None, help="accepted for compatibility; unused")
    parser.add_argument("--precision", type=int, default=2, choices=(1, 2), help="1 single (fdtd_main_gpu_single.x), 2 double")
    parser.add_argument("--energy", action="store_true", help="do energy calc (the reference's balance, evaluated on the device)")
    parser.add_argument("--device", type=int, default=None)
    parser.add_argument("--gpu_folder", action="store_true", help="rotate / fold (FCC) / sort the folder's datasets in memory first, as sim_setup's "
                        "save_folder_gpu does on disk: a plain folder then runs like its gpu folder (half the stored nodes for FCC)")
    parser.add_argument("--equal_slabs", action="store_true", help="multi-GPU: the reference's equal-plane split instead of the cost-weighted one")
    parser.add_argument("--timing", action="store_true", help="also print the reference's 'Air update' / 'Boundary loop' lines (CUDA events around every air launch)")
    # accepted so that the reference's command lines keep working (sim_fdtd.py:899-906); plotting is not part of the simulation step
    parser.add_argument("--plot", action="store_true", help="not available: use SimEngine.gather_slice between run_steps batches")
    parser.add_argument("--draw_backend", type=str, default="matplotlib", help="ignored")
    parser.add_argument("--json_model", type=str, default=None, help="ignored")
    parser.add_argument("--abc", action="store_true", help="ignored (the absorbing shell is always applied, as in the reference)")
    args = parser.parse_args(argv)
    if args.plot:
        parser.error("--plot: live plots are outside the simulation step; SimEngine.gather_slice(ix|iy|iz) returns the cuts")
    if args.data_dir is None:
        args.data_dir = os.getcwd()  # the C binaries run in the data folder (fdtd_main.c:35)
    eng = SimEngine(args.data_dir, energy_on=args.energy, nthreads=args.nthreads, precision=args.precision, device=args.device, timing=args.timing,
                    gpu_folder=args.gpu_folder, balance=not args.equal_slabs)
    eng.load_h5_data()
    eng.setup_mask()
    eng.allocate_mem()
    eng.set_coeffs()
    eng.checks()
    eng.run_all(args.nsteps)
    eng.save_outputs()
    eng.print_last_samples(5)
    if args.energy:
        eng.print_last_energy(5)
    eng.close()


if __name__ == "__main__":
    main()
