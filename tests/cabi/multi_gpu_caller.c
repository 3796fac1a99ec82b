/* A C-only caller of the cntc_* C-ABI (no Python, no torch, no torchrun): a batch of independent contact cases solved once
 * on the current device and once spread over several devices by the library's own scheduler (cb200_set_devices), results
 * compared bit for bit.  This is the call pattern of a multibody code with one result element per wheel
 * (reference: src/test_table.f90:196-292, one result element per OpenMP thread).
 *
 *   gcc -O2 -I include tests/cabi/multi_gpu_caller.c -L contact_b200/lib -lcontact_addon_b200 -Wl,-rpath,... -o caller
 *   ./caller 0,1 64        devices 0 and 1, 64 cases        (./caller 0,0 runs both shards on device 0: same code path)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "contact_addon_b200.h"

#define MX 45
#define MY 41
#define NPOT (MX * MY)

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void setup(int ire, int icase)
{
    int icp = 1, imodul = 3, ifcver, ierr, len = 1;
    cntc_initialize(&ire, &imodul, &ifcver, &ierr, " ", &len);
    int flags[4] = { 1972 /* ic_tang */, 1973 /* ic_norm */, 1974 /* ic_force */, 1983 /* ic_iestim */ }, vals[4] = { 3, 1, 0, 0 }, n4 = 4;
    cntc_setflags(&ire, &icp, &n4, flags, vals);
    int g0 = 0, ni = 4, nr = 1, ip[4] = { 999, 100, 30, 1 }; double rp[1] = { 1e-6 };
    cntc_setsolverflags(&ire, &icp, &g0, &ni, ip, &nr, rp);
    double mat[4] = { 0.28, 0.28, 82000., 82000. }; int m0 = 0;
    cntc_setmaterialparameters(&ire, &icp, &m0, &n4, mat);
    double fr[2] = { 0.3, 0.3 }; int n2 = 2;
    cntc_setfrictionmethod(&ire, &icp, &m0, &n2, fr);
    double dx = 0.1, pot[6] = { MX, MY, -0.5 * MX * 0.1, -0.5 * MY * 0.1, 0.1, 0.1 }; int ipot = 1, n6 = 6;
    cntc_setpotcontact(&ire, &icp, &ipot, &n6, pot);
    double b[6] = { 0.004116, 0, 0.006, 0, 0, 0 }; int ibase = 1;
    cntc_setundeformeddistc(&ire, &icp, &ibase, &n6, b);
    double chi = 0.0;
    cntc_setrollingstepsize(&ire, &icp, &chi, &dx);
    /* a deterministic sweep of loads and creepages */
    double u = (double) ((icase * 37) % 101) / 100.0 - 0.5, w = (double) ((icase * 53) % 97) / 96.0 - 0.5;
    double fn = 4000.0 * (1.0 + 0.4 * u), cksi = 2e-3 * w, ceta = 1e-3 * u, cphi = 3e-4 * (u - w);
    cntc_setnormalforce(&ire, &icp, &fn);
    cntc_setcreepages(&ire, &icp, &cksi, &ceta, &cphi);
}

int main(int argc, char **argv)
{
    int devs[16], nd = 0, ncase = argc > 2 ? atoi(argv[2]) : 64;
    char *list = strdup(argc > 1 ? argv[1] : "0,0");
    for (char *t = strtok(list, ","); t && nd < 16; t = strtok(NULL, ",")) devs[nd++] = atoi(t);
    if (ncase < 16 || ncase > 999) { fprintf(stderr, "ncase must be 16..999\n"); return 2; }
    int *ire = malloc(sizeof(int) * ncase), *ierr = malloc(sizeof(int) * ncase), icp = 1, npot = NPOT;
    double *ref = malloc(sizeof(double) * 3 * NPOT * (size_t) ncase), *pn = malloc(sizeof(double) * 3 * NPOT);
    int *elref = malloc(sizeof(int) * NPOT * (size_t) ncase), *el = malloc(sizeof(int) * NPOT);
    double wall[2] = { 0, 0 };
    long nbad = 0;
    for (int pass = 0; pass < 2; pass++) {
        int used = cb200_set_devices(pass == 0 ? 1 : nd, pass == 0 ? NULL : devs);
        if (used < 0) { fprintf(stderr, "cb200_set_devices: %s\n", cb200_last_error()); return 3; }
        for (int rep = 0; rep < 2; rep++) {                       /* second repetition: coefficient caches warm on every device */
            for (int k = 0; k < ncase; k++) { ire[k] = k + 1; setup(ire[k], k); }
            double t0 = now();
            cntc_calculate_batch(&ncase, ire, &icp, ierr);
            wall[pass] = now() - t0;
            for (int k = 0; k < ncase; k++) if (ierr[k] < 0) { fprintf(stderr, "case %d: ierror %d (%s)\n", k, ierr[k], cb200_last_error()); return 4; }
        }
        for (int k = 0; k < ncase; k++) {
            cntc_gettractions(&ire[k], &icp, &npot, pn, pn + NPOT, pn + 2 * NPOT);
            cntc_getelementdivision(&ire[k], &icp, &npot, el);
            if (pass == 0) { memcpy(ref + 3 * NPOT * (size_t) k, pn, sizeof(double) * 3 * NPOT); memcpy(elref + NPOT * (size_t) k, el, sizeof(int) * NPOT); }
            else {
                if (memcmp(ref + 3 * NPOT * (size_t) k, pn, sizeof(double) * 3 * NPOT) != 0) nbad++;
                if (memcmp(elref + NPOT * (size_t) k, el, sizeof(int) * NPOT) != 0) nbad++;
            }
        }
        printf("pass %d: %d device(s) in use, %d cases in %.3f s (%.1f cases/s)\n", pass, used, ncase, wall[pass], ncase / wall[pass]);
    }
    int ncon = 0;
    for (int i = 0; i < NPOT; i++) ncon += elref[i] >= 1;
    printf("case 0: %d elements in contact; %ld of %d result arrays differ between the passes\n", ncon, nbad, 2 * ncase);
    cntc_finalizelast();
    return (nbad == 0 && ncon > 100) ? 0 : 1;
}
