"""Synthetic workloads named by BASELINE.json / SURVEY.md section 8(d).

Geometries are the reference's own test molecules (tests/Test.py:29-76, bohr);
the cluster and the perturbed-batch generators follow SURVEY.md 8(d) literally
(seeds 20260101 / 20260102) so that every process builds bit-identical inputs.
"""
import numpy as np

H2O = [(8, (0.0, 0.0, 0.091685801102911746)),
       (1, (1.4229678834888837, 0.0, -0.98120954931681137)),
       (1, (-1.4229678834888837, 0.0, -0.98120954931681137))]

CH4 = [(6, (0.0, 0.0, 0.0)),
       (1, (1.182181057825485, -1.182181057825485, 1.182181057825485)),
       (1, (-1.182181057825485, 1.182181057825485, 1.182181057825485)),
       (1, (1.182181057825485, 1.182181057825485, -1.182181057825485)),
       (1, (-1.182181057825485, -1.182181057825485, -1.182181057825485))]

H2 = [(1, (0.0, 0.0, 0.6921792096923653)), (1, (0.0, 0.0, -0.6921792096923653))]

H2_EXAMPLE = [(1, (0.0, 0.0, 0.20165898)), (1, (0.0, 0.0, -1.64601435))]

HCN = [(1, (0.0, 0.0, 0.0)), (6, (0.0, 0.0, 2.0125581778365533)), (7, (0.0, 0.0, 4.1914122426680516))]

CH2O = [(8, (0.0, 0.0, 1.33648844)), (6, (0.0, 0.0, -0.97135846)),
        (1, (0.0, 1.76607674, -2.17632039)), (1, (0.0, -1.76607674, -2.17632039))]


def _random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    a, b, c, d = q
    return np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                     [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                     [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]])


def water_cluster(n_mol=32, spacing=5.67, jitter=0.05, seed=20260101):
    """n_mol rigidly rotated, jittered copies of the H2O monomer on a 4x4x2-style grid.

    SURVEY.md 8(d) config 4: 32 copies, N(0, 0.05 bohr) jitter per atom, numpy default_rng(20260101).
    Spacing: SURVEY.md 8(d) specifies 5.67 bohr (3.0 A), the default here.  From the CORE guess the reference's plain
    Roothaan iteration (no damping / DIIS, Energy.py:159-195) does not converge such a dense cluster: the core-Hamiltonian
    guess piles the electrons onto the central monomers (unscreened attraction of all nuclei) and the iteration falls into
    a 2-cycle, i.e. the reference returns its 99999 sentinel.  Measured with the CPU oracle (= the reference's behaviour):
    n=8 converges from 7 bohr on, n=16 converges at 12, 16 and 20 bohr (12 iterations) and fails at <= 10 bohr, n=32
    fails at 12 and 16 bohr and converges at 24 bohr (13 iterations, E = -2398.5489 Eh).  The reference's own remedy is
    `readguess` (Energy.py:148-157): started from the superposition of monomer SCFs (`monomer_guess` below) the dense
    n=32 cluster converges in 8 iterations, E = -2398.52725184 Eh (n=8: 7 iterations, E = -599.6434204542 Eh).
    Returns the `mol` list System_mol takes; ne = 10*n_mol.
    """
    rng = np.random.default_rng(seed)
    mono = np.array([p for _, p in H2O])
    mono = mono - mono.mean(axis=0)
    grid = [(ix, iy, iz) for iz in range(64) for iy in range(4) for ix in range(4)][:n_mol]
    mol = []
    for g in grid:
        rot = _random_rotation(rng)
        pos = mono @ rot.T + np.array(g, dtype=float) * spacing + rng.normal(scale=jitter, size=(3, 3))
        for (z, _), p in zip(H2O, pos):
            mol.append((z, (float(p[0]), float(p[1]), float(p[2]))))
    return mol


def perturbed_methanes(n=1024, sigma=0.05, seed=20260102):
    """n copies of CH4 with i.i.d. N(0, sigma) on all 15 coordinates."""
    rng = np.random.default_rng(seed)
    base = np.array([p for _, p in CH4])
    out = []
    for _ in range(n):
        pos = base + rng.normal(scale=sigma, size=base.shape)
        out.append([(z, (float(p[0]), float(p[1]), float(p[2]))) for (z, _), p in zip(CH4, pos)])
    return out


def monomer_guess(mol, basis, solve=None, atoms_per_monomer=3, electrons_per_monomer=10, device=0):
    """Block-diagonal `readguess` matrix (AO x MO, Energy.py:148-157) for a cluster of identical monomers: every monomer's
    own RHF coefficients, the occupied columns of all monomers first (the reference builds D0 from the first ne columns,
    Energy.py:150-156), the virtual ones after them.  This is what a reference user would assemble from `printguess`
    files of the monomer runs.  `solve(System_mol) -> C` runs one monomer SCF; default: the GPU path (Energy.rhf)."""
    from .Molecule import System_mol
    if solve is None:
        from . import Energy

        def solve(sm):
            r = Energy.rhf(np.log(sm.alpha), sm.coef, sm.xyz, sm.l, sm.charges, sm.atom, sm.list_contr, sm.ne, max_scf=100, device=device)
            if r["status"] != 0:
                raise RuntimeError("monomer SCF did not converge")
            return r["C"]
    nmono = len(mol) // atoms_per_monomer
    blocks = []
    for m in range(nmono):
        sm = System_mol(mol[atoms_per_monomer * m:atoms_per_monomer * (m + 1)], basis, electrons_per_monomer, "monomer")
        blocks.append(np.asarray(solve(sm), dtype=np.float64))
    n = sum(b.shape[0] for b in blocks)
    nocc = electrons_per_monomer // 2
    C0 = np.zeros((n, n))
    row, occ_col, virt_col = 0, 0, nocc * nmono
    for b in blocks:
        k = b.shape[0]
        C0[row:row + k, occ_col:occ_col + nocc] = b[:, :nocc]
        C0[row:row + k, virt_col:virt_col + k - nocc] = b[:, nocc:]
        row += k; occ_col += nocc; virt_col += k - nocc
    return C0
