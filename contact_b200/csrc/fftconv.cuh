// fftconv.cuh -- the influence-coefficient product u = A p as a pruned, zero-padded 2-D FP64 FFT convolution that
// lives entirely in one CTA's shared memory (sm_100a: up to 227 KB per CTA).
//
// Replaces fft_VecAijPj of the reference (/root/reference/src/m_aijpj.f90:712-1015): pad p(mx,my) into a
// 2Fx x 2Fy array (Fx = opt_fft_size(mx)), real 2-D FFT, pointwise multiply with the transformed coefficients,
// inverse FFT, read the result at offset (Fx,Fy), store where the element-division mask selects.
//
// B200-first formulation (not MKL's):
//   * row pass    : the my non-zero rows only; each real row of length 2Fx is transformed as a packed complex FFT of
//                   length Fx (+ split step) -> S[kx][iy], kx = 0..Fx, batch index iy fastest (bank-conflict free).
//   * column pass : chunks of C columns kx; zero-padded complex FFT of length 2Fy in W[j][c] (c fastest), multiply by
//                   C^ in the last stage's registers, inverse FFT, keep rows Fy..Fy+my-1 only.
//   * inverse rows: Hermitian merge, packed inverse FFT of length Fx, keep columns Fx..Fx+mx-1 only.
//   All FFT stages are in place: decimation in frequency forward, decimation in time backward, so the spectrum
//   stays in digit-reversed order and C^ is simply stored in that same order (built by the same code) -- no
//   reordering pass, no second buffer.  1/(4 Fx Fy) and 1/G are folded into C^.
//
// Every phase is a plain function of (tid, nthr) so that the identical code can be stepped through on the host
// (tests/host_emul) for validation without a GPU; on the device phases are separated by __syncthreads().
#pragma once
#include "fft_radix.cuh"

namespace cb200 {

#define CB_MAXSTAGE 8

struct ConvPlan {
    int mx, my, npot;
    int Fx, Fy;                 // half sizes of the padded array (opt_fft_size)
    int Lx, Ly;                 // transform lengths: Lx = Fx (packed real rows), Ly = 2 Fy
    int nsx, nsy;               // number of stages
    int rx[CB_MAXSTAGE], ry[CB_MAXSTAGE];   // radices, DIF order
    int SY;                     // batch stride of S (>= number of rows held, odd)
    int C;                      // columns per chunk
    int nchunk;                 // ceil((Fx+1)/C)
    int chat_len;               // nchunk * Ly * C  (cd elements per coefficient block)
    // byte offsets into dynamic shared memory
    int off_S, off_W, off_twx, off_twy, off_posx, off_red, smem_bytes;
    // device tables
    const cd *twx;              // [2Fx] exp(-2 pi i k / 2Fx)
    const cd *twy;              // [2Fy] exp(-2 pi i k / 2Fy)
    const unsigned short *posx; // [Lx] position of frequency k after the DIF stages
};

// ------------------------------------------------------------------------------------------------------------
// generic in-place stage over a batch of transforms: element e of transform c lives at buf[e*estride + c]
// ------------------------------------------------------------------------------------------------------------
template <int R, bool INV>
CB_HD void fft_stage(cd *buf, int nbatch, int estride, int L, int ns, const cd *tw, int twmul, int tid, int nthr)
{
    const int m = ns / R;
    const int items = (L / R) * nbatch;
    const int tstep = (L / ns) * twmul;
    for (int w = tid; w < items; w += nthr) {
        const int c = w % nbatch, g = w / nbatch;
        const int j = g % m, blk = g / m;
        cd *p = buf + (size_t) (blk * ns + j) * estride + c;
        const size_t qs = (size_t) m * estride;
        cd x[R];
#pragma unroll
        for (int q = 0; q < R; q++) x[q] = p[q * qs];
        if (INV && m > 1) {
            const int t = tstep * j;
#pragma unroll
            for (int q = 1; q < R; q++) x[q] = cmulc(x[q], tw[t * q]);
        }
        Dft<R, INV>::run(x);
        if (!INV && m > 1) {
            const int t = tstep * j;
#pragma unroll
            for (int q = 1; q < R; q++) x[q] = cmul(x[q], tw[t * q]);
        }
#pragma unroll
        for (int q = 0; q < R; q++) p[q * qs] = x[q];
    }
}

template <bool INV>
CB_HD void fft_stage_r(int r, cd *buf, int nbatch, int estride, int L, int ns, const cd *tw, int twmul, int tid, int nthr)
{
    switch (r) {
    case 2:  fft_stage<2, INV>(buf, nbatch, estride, L, ns, tw, twmul, tid, nthr); break;
    case 3:  fft_stage<3, INV>(buf, nbatch, estride, L, ns, tw, twmul, tid, nthr); break;
    case 4:  fft_stage<4, INV>(buf, nbatch, estride, L, ns, tw, twmul, tid, nthr); break;
    case 5:  fft_stage<5, INV>(buf, nbatch, estride, L, ns, tw, twmul, tid, nthr); break;
    case 7:  fft_stage<7, INV>(buf, nbatch, estride, L, ns, tw, twmul, tid, nthr); break;
    case 8:  fft_stage<8, INV>(buf, nbatch, estride, L, ns, tw, twmul, tid, nthr); break;
    case 9:  fft_stage<9, INV>(buf, nbatch, estride, L, ns, tw, twmul, tid, nthr); break;
    case 16: fft_stage<16, INV>(buf, nbatch, estride, L, ns, tw, twmul, tid, nthr); break;
    default: break;
    }
}

// last forward stage (m = 1, no twiddles) + pointwise multiply with C^ + first inverse stage, in registers
template <int R>
CB_HD void fft_stage_mid(cd *buf, int nbatch, int estride, int L, const cd *chat, int tid, int nthr)
{
    const int items = (L / R) * nbatch;
    for (int w = tid; w < items; w += nthr) {
        const int c = w % nbatch, g = w / nbatch;
        const size_t o = (size_t) (g * R) * estride + c;
        cd x[R];
#pragma unroll
        for (int q = 0; q < R; q++) x[q] = buf[o + (size_t) q * estride];
        Dft<R, false>::run(x);
#pragma unroll
        for (int q = 0; q < R; q++) x[q] = cmul(x[q], chat[o + (size_t) q * estride]);
        Dft<R, true>::run(x);
#pragma unroll
        for (int q = 0; q < R; q++) buf[o + (size_t) q * estride] = x[q];
    }
}

CB_HD void fft_stage_mid_r(int r, cd *buf, int nbatch, int estride, int L, const cd *chat, int tid, int nthr)
{
    switch (r) {
    case 2:  fft_stage_mid<2>(buf, nbatch, estride, L, chat, tid, nthr); break;
    case 3:  fft_stage_mid<3>(buf, nbatch, estride, L, chat, tid, nthr); break;
    case 4:  fft_stage_mid<4>(buf, nbatch, estride, L, chat, tid, nthr); break;
    case 5:  fft_stage_mid<5>(buf, nbatch, estride, L, chat, tid, nthr); break;
    case 7:  fft_stage_mid<7>(buf, nbatch, estride, L, chat, tid, nthr); break;
    case 8:  fft_stage_mid<8>(buf, nbatch, estride, L, chat, tid, nthr); break;
    case 9:  fft_stage_mid<9>(buf, nbatch, estride, L, chat, tid, nthr); break;
    case 16: fft_stage_mid<16>(buf, nbatch, estride, L, chat, tid, nthr); break;
    default: break;
    }
}

// ------------------------------------------------------------------------------------------------------------
// row pass pieces
// ------------------------------------------------------------------------------------------------------------

// Source of real rows: either a traction column p(mx,my) placed at the low corner (m_aijpj.f90:932-939), or an
// influence-coefficient block cf(-cmx:cmx-1,-cmy:cmy-1) placed with cf(0,0) at (Fx,Fy) (m_aijpj.f90:894-901).
struct RowSrc {
    const double *base;
    int kind;            // 0: tractions, 1: coefficients
    int mx, my;          // tractions: grid size ; coefficients: limits min(F, m) used by the reference copy loop
    int cmx, cmy;        // coefficients: allocated half sizes of cf
    int Fx, Fy;
    int row0;            // first padded row handled in this batch (coefficients are done in slabs of rows)
};

CB_HD double rowsrc_get(const RowSrc &s, int row, int col)
{
    if (s.kind == 0) {
        return (row < s.my && col < s.mx) ? s.base[(size_t) row * s.mx + col] : 0.0;
    } else {
        const int iy = row + s.row0 - s.Fy, ix = col - s.Fx;
        if (iy < -s.my || iy >= s.my || ix < -s.mx || ix >= s.mx) return 0.0;
        return s.base[(size_t) (iy + s.cmy) * (2 * s.cmx) + (ix + s.cmx)];
    }
}

// S[j][b] = x[b][2j] + i x[b][2j+1]  for j < Lx, b < nbatch
CB_HD void row_load(const ConvPlan &P, cd *S, int SY, int nbatch, const RowSrc &src, int tid, int nthr)
{
    const int items = P.Lx * nbatch;
    for (int w = tid; w < items; w += nthr) {
        const int j = w % P.Lx, b = w / P.Lx;
        S[(size_t) j * SY + b] = make_double2(rowsrc_get(src, b, 2 * j), rowsrc_get(src, b, 2 * j + 1));
    }
}

// split step of the packed real transform: Z (scrambled) -> X[k], k = 0..Lx, X[k] at row posx[k] (k<Lx), X[Lx] at row Lx
CB_HD void row_split(const ConvPlan &P, cd *S, int SY, int nbatch, const cd *twx, const unsigned short *posx,
                     int tid, int nthr)
{
    const int L = P.Lx, npair = L / 2 + 1;
    const int items = npair * nbatch;
    for (int w = tid; w < items; w += nthr) {
        const int b = w % nbatch, k = w / nbatch;
        if (k == 0) {
            cd a = S[b];
            S[b] = make_double2(a.x + a.y, 0.0);
            S[(size_t) L * SY + b] = make_double2(a.x - a.y, 0.0);
        } else if (2 * k == L) {
            cd *pa = S + (size_t) posx[k] * SY + b;
            *pa = cconj(*pa);
        } else {
            cd *pa = S + (size_t) posx[k] * SY + b, *pb = S + (size_t) posx[L - k] * SY + b;
            cd a = *pa, bb = *pb;
            cd e = make_double2(0.5 * (a.x + bb.x), 0.5 * (a.y - bb.y));          // (a + conj b)/2
            cd d = make_double2(0.5 * (a.x - bb.x), 0.5 * (a.y + bb.y));          // (a - conj b)/2
            cd t = cmul(twx[k], make_double2(d.y, -d.x));                        // w^k * (-i) d
            *pa = cadd(e, t);
            *pb = cconj(csub(e, t));
        }
    }
}

// merge step of the inverse packed real transform: X -> Z' = 2 Z (scrambled positions)
CB_HD void row_merge(const ConvPlan &P, cd *S, int SY, int nbatch, const cd *twx, const unsigned short *posx,
                     int tid, int nthr)
{
    const int L = P.Lx, npair = L / 2 + 1;
    const int items = npair * nbatch;
    for (int w = tid; w < items; w += nthr) {
        const int b = w % nbatch, k = w / nbatch;
        if (k == 0) {
            cd p = S[b], q = S[(size_t) L * SY + b];
            cd u = make_double2(p.x + q.x, p.y - q.y), d = make_double2(p.x - q.x, p.y + q.y);
            S[b] = make_double2(u.x - d.y, u.y + d.x);                           // u + i d
        } else if (2 * k == L) {
            cd *pa = S + (size_t) posx[k] * SY + b;
            *pa = make_double2(2.0 * pa->x, -2.0 * pa->y);
        } else {
            cd *pa = S + (size_t) posx[k] * SY + b, *pb = S + (size_t) posx[L - k] * SY + b;
            cd p = *pa, q = *pb;
            cd u = make_double2(p.x + q.x, p.y - q.y);                           // p + conj q
            cd d = make_double2(p.x - q.x, p.y + q.y);                           // p - conj q
            cd v = cmulc(make_double2(-d.y, d.x), twx[k]);                       // i d conj(w^k)
            *pa = cadd(u, v);
            *pb = cconj(csub(u, v));
        }
    }
}

// store the wanted part of the result: u(ix,iy) = x[iy][Fx+ix], ix < mx  (m_aijpj.f90:978-1005)
// mask_mode 0: all elements (AllElm), 1: el >= Adhes (AllInt).  add: u += result.
CB_HD void row_store(const ConvPlan &P, const cd *S, int SY, double *u, const int *el, int mask_mode, int add,
                     int tid, int nthr)
{
    for (int ii = tid; ii < P.npot; ii += nthr) {
        const int ix = ii % P.mx, iy = ii / P.mx;
        if (mask_mode == 1 && el[ii] < 1) continue;
        const int xi = P.Fx + ix;
        const cd z = S[(size_t) (xi >> 1) * SY + iy];
        const double v = (xi & 1) ? z.y : z.x;
        u[ii] = add ? u[ii] + v : v;
    }
}

// ------------------------------------------------------------------------------------------------------------
// column pass pieces.  Chunk ch holds S rows r = ch*C + c, c < C (rows beyond Fx are treated as zero).
// ------------------------------------------------------------------------------------------------------------
CB_HD void col_load(const ConvPlan &P, const cd *S, int SY, int n_in, cd *W, int ch, int tid, int nthr)
{
    const int items = P.Ly * P.C;
    for (int w = tid; w < items; w += nthr) {
        const int c = w % P.C, j = w / P.C;
        const int r = ch * P.C + c;
        W[w] = (j < n_in && r <= P.Fx) ? S[(size_t) r * SY + j] : make_double2(0.0, 0.0);
    }
}

CB_HD void col_store(const ConvPlan &P, cd *S, int SY, const cd *W, int ch, int tid, int nthr)
{
    const int items = P.my * P.C;
    for (int w = tid; w < items; w += nthr) {
        const int c = w % P.C, iy = w / P.C;
        const int r = ch * P.C + c;
        if (r <= P.Fx) S[(size_t) r * SY + iy] = W[(size_t) (P.Fy + iy) * P.C + c];
    }
}

// coefficient builder: dump the fully forward-transformed chunk, scaled, as C^
CB_HD void col_dump(const ConvPlan &P, const cd *W, cd *chat, int ch, double scale, int tid, int nthr)
{
    const int items = P.Ly * P.C;
    for (int w = tid; w < items; w += nthr)
        chat[(size_t) ch * items + w] = cscale(W[w], scale);
}

}  // namespace cb200
