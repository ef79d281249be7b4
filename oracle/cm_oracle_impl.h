/*
 * cm_oracle_impl.h -- body of the CPU oracle, compiled once per precision.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * contrast-maximization hot path; it is the checker the CUDA kernels are
 * compared against, never the thing that is shipped or measured (only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it).
 *
 * The includer defines:
 *   REAL        float | double
 *   FN(name)    name##_f32 | name##_f64
 *   R_FMA, R_FLOOR, R_FABS, R_RINT   the libm functions of that precision
 *
 * Every function cites the reference lines it restates (paths relative to the
 * upstream repository tudelft/taming_event_flow).  Arithmetic order follows
 * the reference's eager CPU PyTorch path operation by operation; the build
 * uses -ffp-contract=off so that the only fused multiply-adds are the explicit
 * ones (ATen's CPU grid_sample accumulates its four taps as an FMA chain --
 * established bit-exactly in tests/test_primitives_vs_reference_live.py and tests/test_oracle_vs_reference_live.py).
 */

/* ------------------------------------------------------------------------ */
/* utils/iwe.py:17-40 get_event_flow + ATen grid_sampler_2d (bilinear,      */
/* zeros padding, align_corners=True).  y,x are pixel coordinates.          */
/* taps[] (optional) receives the 4 tap values of both maps and the weights */
/* so the backward can reuse them.                                          */
/* ------------------------------------------------------------------------ */
typedef struct {
    int y0, x0;          /* north-west tap                                  */
    int ok[4];           /* tap in bounds: nw, ne, sw, se                   */
    REAL w[4];           /* nw, ne, sw, se interpolation weights            */
    REAL ax, ay;         /* fractional offsets (w and n in ATen's naming)   */
    REAL vx[4], vy[4];   /* tap values of the x map / y map (0 if outside)  */
} FN(taps_t);

static inline void FN(unnormalised)(REAL y, REAL x, int H, int W, REAL *iy, REAL *ix)
{
    /* utils/iwe.py:30-31 : 2 * v / (size - 1) - 1, evaluated left to right */
    REAL gy = ((REAL)2 * y) / (REAL)(H - 1) - (REAL)1;
    REAL gx = ((REAL)2 * x) / (REAL)(W - 1) - (REAL)1;
    /* ATen align_corners=True: (g + 1) * ((size - 1) / 2)                  */
    *iy = (gy + (REAL)1) * ((REAL)(H - 1) / (REAL)2);
    *ix = (gx + (REAL)1) * ((REAL)(W - 1) / (REAL)2);
}

static inline void FN(sample_flow)(const REAL *mapx, const REAL *mapy, int H, int W,
                                   REAL y, REAL x, REAL *vy, REAL *vx, FN(taps_t) *tp)
{
    REAL iy, ix;
    FN(unnormalised)(y, x, H, W, &iy, &ix);
    REAL fx0 = R_FLOOR(ix), fy0 = R_FLOOR(iy);
    REAL w_ = ix - fx0, e_ = (REAL)1 - w_;
    REAL n_ = iy - fy0, s_ = (REAL)1 - n_;
    FN(taps_t) t;
    t.w[0] = s_ * e_; t.w[1] = s_ * w_; t.w[2] = n_ * e_; t.w[3] = n_ * w_;
    t.ax = w_; t.ay = n_;
    /* positions on the chain are always inside the image; stand-alone calls
       may be anywhere, so clamp before the integer conversion              */
    if (!(fx0 >= (REAL)-2 && fx0 <= (REAL)(W + 1) && fy0 >= (REAL)-2 && fy0 <= (REAL)(H + 1))) {
        t.y0 = -2; t.x0 = -2;
    } else {
        t.y0 = (int)fy0; t.x0 = (int)fx0;
    }
    const int ty[4] = { t.y0, t.y0, t.y0 + 1, t.y0 + 1 };
    const int tx[4] = { t.x0, t.x0 + 1, t.x0, t.x0 + 1 };
    for (int k = 0; k < 4; ++k) {
        t.ok[k] = (ty[k] >= 0 && ty[k] < H && tx[k] >= 0 && tx[k] < W);
        t.vx[k] = t.ok[k] ? mapx[(size_t)ty[k] * W + tx[k]] : (REAL)0;
        t.vy[k] = t.ok[k] ? mapy[(size_t)ty[k] * W + tx[k]] : (REAL)0;
    }
    /* ATen CPU kernel: nw*w_nw, then three fused multiply-adds             */
    REAL ox = t.vx[0] * t.w[0], oy = t.vy[0] * t.w[0];
    for (int k = 1; k < 4; ++k) { ox = R_FMA(t.vx[k], t.w[k], ox); oy = R_FMA(t.vy[k], t.w[k], oy); }
    *vx = ox; *vy = oy;
    if (tp) *tp = t;
}

/* utils/iwe.py:43-60 purge_unfeasible: inclusive bounds [0, res-1]         */
static inline int FN(inside)(REAL y, REAL x, int H, int W)
{
    return (y >= (REAL)0) && (y <= (REAL)H - (REAL)1) && (x >= (REAL)0) && (x <= (REAL)W - (REAL)1);
}

/* ------------------------------------------------------------------------ */
/* utils/iwe.py:63-113 get_interpolation (bilinear branch), one event.      */
/* corner order TL,TR,BL,BR (utils/iwe.py:90-94).  cy/cx are the corner     */
/* coordinates as REAL (before masking), ok = strict in-image test (:103).  */
/* wy/wx are the clamped 1-D weights (:100), w = wy*wx*ok (:107).           */
/* ------------------------------------------------------------------------ */
typedef struct {
    REAL cy[4], cx[4], wy[4], wx[4], w[4];
    int ok[4];
    long pix[4];
} FN(corners_t);

static inline void FN(corners)(REAL y, REAL x, int H, int W, FN(corners_t) *c)
{
    REAL top = R_FLOOR(y), bot = R_FLOOR(y + (REAL)1);
    REAL left = R_FLOOR(x), right = R_FLOOR(x + (REAL)1);
    const REAL cy[4] = { top, top, bot, bot };
    const REAL cx[4] = { left, right, left, right };
    for (int k = 0; k < 4; ++k) {
        c->cy[k] = cy[k]; c->cx[k] = cx[k];
        REAL uy = (REAL)1 - R_FABS(y - cy[k]);
        REAL ux = (REAL)1 - R_FABS(x - cx[k]);
        c->wy[k] = uy > (REAL)0 ? uy : (REAL)0;
        c->wx[k] = ux > (REAL)0 ? ux : (REAL)0;
        c->ok[k] = (cy[k] >= (REAL)0) && (cy[k] < (REAL)H) && (cx[k] >= (REAL)0) && (cx[k] < (REAL)W);
        c->w[k] = c->ok[k] ? c->wy[k] * c->wx[k] : (REAL)0;
        c->pix[k] = c->ok[k] ? (long)cy[k] * W + (long)cx[k] : 0;
    }
}

/* derivative of max(0, 1-|v-c|) w.r.t. v under autograd's conventions:
   abs'(0) = 0, max() ties split 1/2               */
static inline REAL FN(d1)(REAL v, REAL c)
{
    REAL d = v - c;
    REAL u = (REAL)1 - R_FABS(d);
    REAL sg = (d > (REAL)0) ? (REAL)1 : ((d < (REAL)0) ? (REAL)-1 : (REAL)0);
    if (u > (REAL)0) return -sg;
    if (u == (REAL)0) return (REAL)-0.5 * sg;
    return (REAL)0;
}

/* ------------------------------------------------------------------------ */
/* slot table: one image set per (scale, sub-window, tref)                  */
/* loss/flow.py:657-668,684-686,730-736                                     */
/* ------------------------------------------------------------------------ */
#ifndef ORC_SLOT_T
#define ORC_SLOT_T
typedef struct { int s, lo, hi, tref, delta, low_tref, high_tref; double coef; } orc_slot;
#endif

static int FN(build_slots)(const orc_cfg *c, orc_slot *out)
{
    int n = 0;
    for (int s = 0; s < c->S; ++s) {
        int L = c->P >> s;
        int delta = (c->mode == 1) ? L : (c->mode == 2 ? L / 2 : L / 4);
        for (int w = 0; w < (1 << s); ++w) {
            int lo = w * L, hi = (w + 1) * L;
            int low_tref = lo, high_tref = hi + 1;
            if (c->mode == 4) { low_tref = lo + delta; high_tref = lo + 3 * delta + 1; }
            for (int tr = low_tref; tr < high_tref; ++tr) {
                if (out) {
                    out[n].s = s; out[n].lo = lo; out[n].hi = hi; out[n].tref = tr; out[n].delta = delta;
                    out[n].low_tref = low_tref; out[n].high_tref = high_tref;
                    out[n].coef = 1.0 / (double)(1 << s) / (double)(2 * delta + 1) / (double)c->S / (double)c->F;
                }
                ++n;
            }
        }
    }
    return n;
}

/* gradient of the total loss w.r.t. one focus_loss value, divided in the    */
/* order autograd unwinds loss/flow.py:730-736 (Linear: :396-402)           */
static inline REAL FN(upstream)(const orc_cfg *c, const orc_slot *q, int linear)
{
    REAL g = (REAL)1;
    g = g / (REAL)c->F;
    g = g / (REAL)c->S;
    g = g / (linear ? (REAL)2 : (REAL)(2 * q->delta + 1));
    g = g / (REAL)(1 << q->s);
    return g;
}

/* ------------------------------------------------------------------------ */
/* Iterative warping chain of one event (loss/flow.py:492-586).             */
/* node[tau] = position at reference time tau, alive[tau] = cumulative      */
/* in-image product along the chain that reaches tau.                       */
/* flow maps of one flow scale: maps[(p*B + b)*2 + ch][H][W], ch0=x, ch1=y  */
/* ------------------------------------------------------------------------ */
static void FN(chain)(const REAL *maps, int B, int H, int W, int P, int b, int t,
                      REAL ts, REAL y0, REAL x0, REAL *ny, REAL *nx, unsigned char *alive)
{
    const size_t HW = (size_t)H * W;
    /* forward: sampling map t+k, landing on node t+k+1 (loss/flow.py:505-510) */
    REAL y = y0, x = x0, tprev = ts; int al = 1;
    for (int k = 0; t + k < P; ++k) {
        const REAL *mx = maps + ((size_t)(t + k) * B + b) * 2 * HW, *my = mx + HW;
        REAL vy, vx; FN(sample_flow)(mx, my, H, W, y, x, &vy, &vx, 0);
        REAL tw = (REAL)(t + k + 1);
        REAL dt = tw - tprev;                 /* utils/iwe.py:14            */
        y = y + dt * vy; x = x + dt * vx; tprev = tw;
        int in = FN(inside)(y, x, H, W);
        y = y * (REAL)in; x = x * (REAL)in; al &= in;       /* utils/iwe.py:58-59 */
        ny[t + k + 1] = y; nx[t + k + 1] = x; alive[t + k + 1] = (unsigned char)al;
    }
    /* backward: sampling map t+k (k=0,-1,..), landing on node t+k (:512-514) */
    y = y0; x = x0; tprev = ts; al = 1;
    for (int k = 0; t + k >= 0; --k) {
        const REAL *mx = maps + ((size_t)(t + k) * B + b) * 2 * HW, *my = mx + HW;
        REAL vy, vx; FN(sample_flow)(mx, my, H, W, y, x, &vy, &vx, 0);
        REAL tw = (REAL)(t + k);
        REAL dt = tw - tprev;
        y = y + dt * vy; x = x + dt * vx; tprev = tw;
        int in = FN(inside)(y, x, H, W);
        y = y * (REAL)in; x = x * (REAL)in; al &= in;
        ny[t + k] = y; nx[t + k] = x; alive[t + k] = (unsigned char)al;
    }
}

/* event set after `update` (loss/flow.py:443-476)                          */
typedef struct {
    long E;              /* total rows over all passes                      */
    const REAL *ev;      /* [E][4] (ts, y, x, p) raw, ts in [0,1]           */
    const REAL *mk;      /* [E][2] (pos, neg)                               */
    const int *n;        /* rows per sample in pass t                       */
    long off[ORC_MAX_P + 1];
    REAL *ts;            /* [E] ts + pass (or round_ts override)            */
    int *pass; int *bat; /* [E] pass index and batch index of each row      */
} FN(evset);

static int FN(evset_init)(FN(evset) *s, const orc_cfg *c, const REAL *ev, const REAL *mk, const int *n)
{
    s->ev = ev; s->mk = mk; s->n = n; s->off[0] = 0;
    for (int t = 0; t < c->P; ++t) s->off[t + 1] = s->off[t] + (long)c->B * n[t];
    s->E = s->off[c->P];
    s->ts = (REAL *)malloc(sizeof(REAL) * (size_t)(s->E > 0 ? s->E : 1));
    s->pass = (int *)malloc(sizeof(int) * (size_t)(s->E > 0 ? s->E : 1));
    s->bat = (int *)malloc(sizeof(int) * (size_t)(s->E > 0 ? s->E : 1));
    for (int t = 0; t < c->P; ++t) {
        long nrow = (long)c->B * n[t];
        REAL mn = 0; int have = 0;
        for (long i = 0; i < nrow; ++i) {
            long e = s->off[t] + i;
            s->ts[e] = ev[e * 4] + (REAL)t;                 /* loss/flow.py:457 */
            s->pass[e] = t; s->bat[e] = (int)(i / n[t]);
            if (!have || s->ts[e] < mn) { mn = s->ts[e]; have = 1; }
        }
        if (c->round_ts) {                                  /* loss/flow.py:461-463 */
            if (!have) return -3;                           /* min() of an empty tensor raises */
            for (long i = 0; i < nrow; ++i) s->ts[s->off[t] + i] = mn + (REAL)0.5;
        }
    }
    return 0;
}
static void FN(evset_free)(FN(evset) *s) { free(s->ts); free(s->pass); free(s->bat); }

/* does window t feed slot q, and is the event alive under q's mask?        */
static inline int FN(slot_takes)(const orc_slot *q, int t)
{
    int lo_e = q->tref - q->delta > q->lo ? q->tref - q->delta : q->lo;      /* loss/flow.py:685 */
    int hi_e = q->tref + q->delta < q->hi ? q->tref + q->delta : q->hi;      /* :686 */
    return t >= lo_e && t < hi_e;
}
static inline int FN(slot_alive)(const orc_cfg *c, const orc_slot *q, const unsigned char *alive /* stride E */, long E, long e)
{
    if (!c->border_comp) return alive[(size_t)q->tref * E + e];              /* loss/flow.py:694 */
    int a = 1;                                                               /* :671-681 */
    for (int tr = q->low_tref; tr < q->high_tref; ++tr) a &= alive[(size_t)tr * E + e];
    return a;
}

/* ------------------------------------------------------------------------ */
/* Iterative loss, forward + analytic backward (loss/flow.py:588-746).      */
/* ------------------------------------------------------------------------ */
int FN(orc_iterative)(const orc_cfg *c,
                      const REAL *flow,                       /* [F][P][B][2][H][W]          */
                      const REAL *ev, const REAL *mk, const int *n_ev,
                      const REAL *dev, const REAL *dmk, const int *n_dev,
                      REAL *loss_out,
                      REAL *gflow,                            /* like flow, or NULL          */
                      REAL *iwe_out,                          /* [F][B][nslots][4][H][W]|NULL*/
                      REAL *nodes_out,                        /* [F][P+1][E][2] | NULL       */
                      unsigned char *alive_out)               /* [F][P+1][E] | NULL          */
{
    if (c->P > ORC_MAX_P || c->P < 1) return -1;
    const int B = c->B, H = c->H, W = c->W, P = c->P, F = c->F;
    const size_t HW = (size_t)H * W;
    orc_slot slots[ORC_MAX_SLOTS];
    int nslots = FN(build_slots)(c, NULL);
    if (nslots > ORC_MAX_SLOTS) return -1;
    FN(build_slots)(c, slots);
    for (int q = 0; q < nslots; ++q) if (slots[q].delta == 0 || slots[q].hi == slots[q].lo) return -2; /* empty torch.cat */
    if (c->mode == 4 && c->border_comp) return -4;   /* reference raises TypeError (None in torch.cat) */

    FN(evset) sets[2];
    int rc;
    if ((rc = FN(evset_init)(&sets[0], c, ev, mk, n_ev))) return rc;
    if ((rc = FN(evset_init)(&sets[1], c, dev, dmk, n_dev))) { FN(evset_free)(&sets[0]); return rc; }

    double loss_total = 0.0;
    if (gflow) memset(gflow, 0, sizeof(REAL) * (size_t)F * P * B * 2 * HW);

    /* chain tables of both sets, reused per flow scale */
    REAL *ny[2], *nx[2]; unsigned char *alive[2];
    for (int k = 0; k < 2; ++k) {
        size_t cells = (size_t)(P + 1) * (size_t)(sets[k].E > 0 ? sets[k].E : 1);
        ny[k] = (REAL *)malloc(sizeof(REAL) * cells); nx[k] = (REAL *)malloc(sizeof(REAL) * cells);
        alive[k] = (unsigned char *)malloc(cells);
    }
    const size_t img_sz = (size_t)B * nslots * 4 * HW;
    REAL *img = (REAL *)malloc(sizeof(REAL) * img_sz);       /* grad+detached sums: cnt+,cnt-,ts+,ts-   */
    ACC *aimg = (ACC *)malloc(sizeof(ACC) * img_sz);         /* partial images of the gradient set ...  */
    ACC *adimg = (ACC *)malloc(sizeof(ACC) * img_sz);        /* ... and of the detached set (ACC = REAL: the reference's sums) */
    double *nnz = (double *)malloc(sizeof(double) * (size_t)B * nslots);
    REAL *gst[2] = { NULL, NULL };                           /* per-node g' of fwd / bwd chain [P+1][E][2] */
    if (gflow) for (int k = 0; k < 2; ++k)
        gst[k] = (REAL *)malloc(sizeof(REAL) * (size_t)(P + 1) * (size_t)(sets[0].E > 0 ? sets[0].E : 1) * 2);

    for (int f = 0; f < F; ++f) {
        const REAL *maps = flow + (size_t)f * P * B * 2 * HW;

        /* ---- phase 1: warp every event to every reference time ---------- */
        for (int k = 0; k < 2; ++k) {
            const FN(evset) *s = &sets[k]; const long E = s->E;
            #pragma omp parallel for schedule(static)
            for (long e = 0; e < E; ++e) {
                REAL ly[ORC_MAX_P + 1], lx[ORC_MAX_P + 1]; unsigned char la[ORC_MAX_P + 1];
                FN(chain)(maps, B, H, W, P, s->bat[e], s->pass[e], s->ts[e], s->ev[e * 4 + 1], s->ev[e * 4 + 2], ly, lx, la);
                for (int tr = 0; tr <= P; ++tr) {
                    ny[k][(size_t)tr * E + e] = ly[tr]; nx[k][(size_t)tr * E + e] = lx[tr]; alive[k][(size_t)tr * E + e] = la[tr];
                }
            }
        }
        if (nodes_out) {
            const long E = sets[0].E;
            for (int tr = 0; tr <= P; ++tr) for (long e = 0; e < E; ++e) {
                nodes_out[(((size_t)f * (P + 1) + tr) * E + e) * 2 + 0] = ny[0][(size_t)tr * E + e];
                nodes_out[(((size_t)f * (P + 1) + tr) * E + e) * 2 + 1] = nx[0][(size_t)tr * E + e];
                if (alive_out) alive_out[((size_t)f * (P + 1) + tr) * E + e] = alive[0][(size_t)tr * E + e];
            }
        }

        /* ---- phase 2: splat (loss/flow.py:81-110, utils/iwe.py:116-136) -- */
        /* order of the fp32 additions follows scatter_add_ on CPU: corner-   */
        /* major, then windows low->high, then rows; grad and detached images */
        /* are built separately and added (loss/flow.py:725-726).             */
        #pragma omp parallel for collapse(2) schedule(dynamic, 1)
        for (int b = 0; b < B; ++b) for (int q = 0; q < nslots; ++q) {
            const orc_slot *sl = &slots[q];
            for (int k = 0; k < 2; ++k) {
                const FN(evset) *s = &sets[k]; const long E = s->E;
                ACC *im = (k == 0 ? aimg : adimg) + ((size_t)b * nslots + q) * 4 * HW;
                memset(im, 0, sizeof(ACC) * 4 * HW);
                for (int corner = 0; corner < 4; ++corner)
                    for (int t = 0; t < P; ++t) {
                        if (!FN(slot_takes)(sl, t)) continue;
                        for (int i = 0; i < s->n[t]; ++i) {
                            long e = s->off[t] + (long)b * s->n[t] + i;
                            REAL al = (REAL)FN(slot_alive)(c, sl, alive[k], E, e);
                            REAL mp = s->mk[e * 2] * al, mn = s->mk[e * 2 + 1] * al;
                            FN(corners_t) cr;
                            FN(corners)(ny[k][(size_t)sl->tref * E + e], nx[k][(size_t)sl->tref * E + e], H, W, &cr);
                            /* loss/flow.py:94-95 */
                            REAL nts = (REAL)1 - R_FABS((REAL)sl->tref - s->ts[e]) / (REAL)sl->delta;
                            REAL w = cr.w[corner], wt = w * nts;
                            long px = cr.pix[corner];
                            im[0 * HW + px] += (ACC)(REAL)(w * mp); im[1 * HW + px] += (ACC)(REAL)(w * mn);
                            im[2 * HW + px] += (ACC)(REAL)(wt * mp); im[3 * HW + px] += (ACC)(REAL)(wt * mn);
                        }
                    }
            }
            REAL *im = img + ((size_t)b * nslots + q) * 4 * HW;
            const ACC *ai = aimg + ((size_t)b * nslots + q) * 4 * HW, *di = adimg + ((size_t)b * nslots + q) * 4 * HW;
            for (size_t i = 0; i < 4 * HW; ++i) im[i] = (REAL)(ai[i] + di[i]);       /* loss/flow.py:725-726 */
        }
        if (iwe_out) {
            for (int b = 0; b < B; ++b)
                memcpy(iwe_out + ((size_t)f * B + b) * nslots * 4 * HW, img + (size_t)b * nslots * 4 * HW, sizeof(REAL) * nslots * 4 * HW);
        }

        /* ---- phase 3: focus loss (loss/flow.py:112-129, :727-736) -------- */
        /* afterwards img holds the gradient images: ch0/1 = dL/dIWE(+,-),    */
        /* ch2/3 = dL/dIWT(+,-) (before the upstream grad, which is 1).       */
        double fl = 0.0;
        #pragma omp parallel for collapse(2) schedule(dynamic, 1) reduction(+ : fl)
        for (int b = 0; b < B; ++b) for (int q = 0; q < nslots; ++q) {
            REAL *im = img + ((size_t)b * nslots + q) * 4 * HW;
            double acc = 0.0; long cnt = 0;
            for (size_t i = 0; i < HW; ++i) {
                REAL ap = im[2 * HW + i] / (im[0 * HW + i] + (REAL)1e-9);
                REAL an = im[3 * HW + i] / (im[1 * HW + i] + (REAL)1e-9);
                acc += (double)(ap * ap) + (double)(an * an);
                cnt += ((im[0 * HW + i] + im[1 * HW + i]) != (REAL)0);
            }
            REAL den = c->loss_scaling ? (REAL)cnt + (REAL)1e-9 : (REAL)1;
            nnz[(size_t)b * nslots + q] = (double)den;
            fl += acc / (double)den * slots[q].coef;
            if (gflow) {
                REAL cf = FN(upstream)(c, &slots[q], 0) / den;
                for (size_t i = 0; i < HW; ++i) {
                    for (int pol = 0; pol < 2; ++pol) {
                        /* autograd's operation order: pow -> grad*(2*A); div -> grad/D and -grad*((T/D)/D) */
                        REAL iw = im[pol * HW + i] + (REAL)1e-9;
                        REAL a = im[(2 + pol) * HW + i] / iw;
                        REAL ga = cf * ((REAL)2 * a);
                        im[(2 + pol) * HW + i] = ga / iw;             /* dL/dIWT */
                        im[pol * HW + i] = -(ga * (a / iw));          /* dL/dIWE */
                    }
                }
            }
        }
        loss_total += fl;

        /* ---- phase 4: backward -------------- */
        if (gflow) {
            const FN(evset) *s = &sets[0]; const long E = s->E;
            REAL *gf = gflow + (size_t)f * P * B * 2 * HW;
            #pragma omp parallel for schedule(static)
            for (long e = 0; e < E; ++e) {
                const int t = s->pass[e], b = s->bat[e];
                REAL gy[ORC_MAX_P + 1], gx[ORC_MAX_P + 1];
                for (int tr = 0; tr <= P; ++tr) { gy[tr] = 0; gx[tr] = 0; }
                /* IWE gradient entering each node */
                for (int q = 0; q < nslots; ++q) {
                    const orc_slot *sl = &slots[q];
                    if (!FN(slot_takes)(sl, t)) continue;
                    REAL al = (REAL)FN(slot_alive)(c, sl, alive[0], E, e);
                    REAL mp = s->mk[e * 2] * al, mn = s->mk[e * 2 + 1] * al;
                    if (mp == (REAL)0 && mn == (REAL)0) continue;
                    const REAL *im = img + ((size_t)b * nslots + q) * 4 * HW;
                    REAL y = ny[0][(size_t)sl->tref * E + e], x = nx[0][(size_t)sl->tref * E + e];
                    REAL nts = (REAL)1 - R_FABS((REAL)sl->tref - s->ts[e]) / (REAL)sl->delta;
                    FN(corners_t) cr; FN(corners)(y, x, H, W, &cr);
                    for (int k = 0; k < 4; ++k) {
                        if (!cr.ok[k]) continue;
                        long px = cr.pix[k];
                        REAL gw = mp * (im[0 * HW + px] + nts * im[2 * HW + px]) + mn * (im[1 * HW + px] + nts * im[3 * HW + px]);
                        gy[sl->tref] += gw * FN(d1)(y, cr.cy[k]) * cr.wx[k];
                        gx[sl->tref] += gw * cr.wy[k] * FN(d1)(x, cr.cx[k]);
                    }
                }
                /* reverse the forward chain: nodes P .. t+1 */
                REAL cy_ = 0, cx_ = 0;
                for (int tr = P; tr >= t + 1; --tr) {
                    REAL in = (REAL)(alive[0][(size_t)tr * E + e] ? 1 : 0);     /* cumulative alive == this step's `in` while upstream alive */
                    /* a dead node blocks everything behind it */
                    REAL gpy = (gy[tr] + cy_) * in, gpx = (gx[tr] + cx_) * in;
                    gst[0][((size_t)tr * E + e) * 2] = gpy; gst[0][((size_t)tr * E + e) * 2 + 1] = gpx;
                    if (gpy == (REAL)0 && gpx == (REAL)0) { cy_ = 0; cx_ = 0; continue; }
                    REAL py = (tr - 1 == t) ? s->ev[e * 4 + 1] : ny[0][(size_t)(tr - 1) * E + e];
                    REAL px = (tr - 1 == t) ? s->ev[e * 4 + 2] : nx[0][(size_t)(tr - 1) * E + e];
                    REAL dt = (tr - 1 == t) ? ((REAL)tr - s->ts[e]) : (REAL)1;
                    const REAL *mx = maps + ((size_t)(tr - 1) * B + b) * 2 * HW, *my = mx + HW;
                    REAL vy, vx; FN(taps_t) tp; FN(sample_flow)(mx, my, H, W, py, px, &vy, &vx, &tp);
                    REAL dvy_dy = ((REAL)1 - tp.ax) * (tp.vy[2] - tp.vy[0]) + tp.ax * (tp.vy[3] - tp.vy[1]);
                    REAL dvy_dx = ((REAL)1 - tp.ay) * (tp.vy[1] - tp.vy[0]) + tp.ay * (tp.vy[3] - tp.vy[2]);
                    REAL dvx_dy = ((REAL)1 - tp.ax) * (tp.vx[2] - tp.vx[0]) + tp.ax * (tp.vx[3] - tp.vx[1]);
                    REAL dvx_dx = ((REAL)1 - tp.ay) * (tp.vx[1] - tp.vx[0]) + tp.ay * (tp.vx[3] - tp.vx[2]);
                    cy_ = gpy + dt * (dvy_dy * gpy + dvx_dy * gpx);
                    cx_ = gpx + dt * (dvy_dx * gpy + dvx_dx * gpx);
                }
                /* reverse the backward chain: nodes 0 .. t */
                cy_ = 0; cx_ = 0;
                for (int tr = 0; tr <= t; ++tr) {
                    REAL in = (REAL)(alive[0][(size_t)tr * E + e] ? 1 : 0);
                    REAL gpy = (gy[tr] + cy_) * in, gpx = (gx[tr] + cx_) * in;
                    gst[1][((size_t)tr * E + e) * 2] = gpy; gst[1][((size_t)tr * E + e) * 2 + 1] = gpx;
                    if (gpy == (REAL)0 && gpx == (REAL)0) { cy_ = 0; cx_ = 0; continue; }
                    REAL py = (tr == t) ? s->ev[e * 4 + 1] : ny[0][(size_t)(tr + 1) * E + e];
                    REAL px = (tr == t) ? s->ev[e * 4 + 2] : nx[0][(size_t)(tr + 1) * E + e];
                    REAL dt = (tr == t) ? ((REAL)tr - s->ts[e]) : (REAL)-1;
                    const REAL *mx = maps + ((size_t)tr * B + b) * 2 * HW, *my = mx + HW;
                    REAL vy, vx; FN(taps_t) tp; FN(sample_flow)(mx, my, H, W, py, px, &vy, &vx, &tp);
                    REAL dvy_dy = ((REAL)1 - tp.ax) * (tp.vy[2] - tp.vy[0]) + tp.ax * (tp.vy[3] - tp.vy[1]);
                    REAL dvy_dx = ((REAL)1 - tp.ay) * (tp.vy[1] - tp.vy[0]) + tp.ay * (tp.vy[3] - tp.vy[2]);
                    REAL dvx_dy = ((REAL)1 - tp.ax) * (tp.vx[2] - tp.vx[0]) + tp.ax * (tp.vx[3] - tp.vx[1]);
                    REAL dvx_dx = ((REAL)1 - tp.ay) * (tp.vx[1] - tp.vx[0]) + tp.ay * (tp.vx[3] - tp.vx[2]);
                    cy_ = gpy + dt * (dvy_dy * gpy + dvx_dy * gpx);
                    cx_ = gpx + dt * (dvy_dx * gpy + dvx_dx * gpx);
                }
            }
            /* flow-map gradient: map p of sample b receives the forward step  */
            /* landing on node p+1 (events of windows <= p) and the backward   */
            /* step landing on node p (windows >= p).  One thread per map, so  */
            /* the additions have a fixed order.                               */
            #pragma omp parallel for collapse(2) schedule(dynamic, 1)
            for (int p = 0; p < P; ++p) for (int b = 0; b < B; ++b) {
                const REAL *mx = maps + ((size_t)p * B + b) * 2 * HW, *my = mx + HW;
                REAL *gmx = gf + ((size_t)p * B + b) * 2 * HW, *gmy = gmx + HW;
                ACC *amx = (ACC *)calloc(2 * HW, sizeof(ACC)), *amy = amx + HW;     /* this map's sums (ACC = REAL: the reference's) */
                for (int dir = 0; dir < 2; ++dir) {
                    int t0 = dir == 0 ? 0 : p, t1 = dir == 0 ? p : P - 1;
                    int node = dir == 0 ? p + 1 : p;
                    for (int t = t0; t <= t1; ++t) for (int i = 0; i < s->n[t]; ++i) {
                        long e = s->off[t] + (long)b * s->n[t] + i;
                        REAL gpy = gst[dir][((size_t)node * E + e) * 2], gpx = gst[dir][((size_t)node * E + e) * 2 + 1];
                        if (gpy == (REAL)0 && gpx == (REAL)0) continue;
                        int src = dir == 0 ? node - 1 : node + 1;   /* node the step started from (t => original location) */
                        int first = dir == 0 ? (node - 1 == t) : (node == t);
                        REAL py = first ? s->ev[e * 4 + 1] : ny[0][(size_t)src * E + e];
                        REAL px = first ? s->ev[e * 4 + 2] : nx[0][(size_t)src * E + e];
                        REAL dt = first ? ((REAL)node - s->ts[e]) : (dir == 0 ? (REAL)1 : (REAL)-1);
                        REAL vy, vx; FN(taps_t) tp; FN(sample_flow)(mx, my, H, W, py, px, &vy, &vx, &tp);
                        const int ty[4] = { tp.y0, tp.y0, tp.y0 + 1, tp.y0 + 1 };
                        const int tx[4] = { tp.x0, tp.x0 + 1, tp.x0, tp.x0 + 1 };
                        for (int k = 0; k < 4; ++k) if (tp.ok[k]) {
                            amy[(size_t)ty[k] * W + tx[k]] += (ACC)(REAL)(dt * tp.w[k] * gpy);
                            amx[(size_t)ty[k] * W + tx[k]] += (ACC)(REAL)(dt * tp.w[k] * gpx);
                        }
                    }
                }
                for (size_t i = 0; i < HW; ++i) { gmx[i] = (REAL)amx[i]; gmy[i] = (REAL)amy[i]; }
                free(amx);
            }
        }
    }
    *loss_out = (REAL)loss_total;

    for (int k = 0; k < 2; ++k) { free(ny[k]); free(nx[k]); free(alive[k]); if (gst[k]) free(gst[k]); }
    free(img); free(aimg); free(adimg); free(nnz);
    FN(evset_free)(&sets[0]); FN(evset_free)(&sets[1]);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Linear loss, forward + analytic backward (loss/flow.py:216-412).         */
/* Slots are (scale, sub-window, end) with end 0 = forward (tref = hi) and  */
/* end 1 = backward (tref = lo); image layout [F][B][nslots][4][H][W].      */
/* ------------------------------------------------------------------------ */
static int FN(linear_slots)(const orc_cfg *c, orc_slot *out)
{
    int n = 0;
    for (int s = 0; s < c->S; ++s) {
        int L = c->P >> s;
        for (int w = 0; w < (1 << s); ++w) for (int end = 0; end < 2; ++end) {
            if (out) {
                out[n].s = s; out[n].lo = w * L; out[n].hi = (w + 1) * L; out[n].tref = end == 0 ? (w + 1) * L : w * L;
                out[n].delta = L; out[n].low_tref = end; out[n].high_tref = 0;
                out[n].coef = 1.0 / (double)(1 << s) / 2.0 / (double)c->S / (double)c->F;
            }
            ++n;
        }
    }
    return n;
}

/* position of one event at one end of its sub-window (loss/flow.py:337-343) */
static inline int FN(linear_pos)(const orc_cfg *c, const orc_slot *q, REAL ts, REAL y0, REAL x0, REAL vy, REAL vx,
                                 REAL *py, REAL *px, REAL *dt_out)
{
    REAL dth = (REAL)q->hi - ts, dtl = (REAL)q->lo - ts;
    REAL fy = y0 + dth * vy, fx = x0 + dth * vx;
    REAL by = y0 + dtl * vy, bx = x0 + dtl * vx;
    int alive = 1;
    if (c->border_comp) {
        int inf = FN(inside)(fy, fx, c->H, c->W), inb = FN(inside)(by, bx, c->H, c->W);
        fy = fy * (REAL)inf; fx = fx * (REAL)inf; by = by * (REAL)inb; bx = bx * (REAL)inb;
        alive = inf & inb;
    }
    if (q->low_tref == 0) { *py = fy; *px = fx; *dt_out = dth; } else { *py = by; *px = bx; *dt_out = dtl; }
    return alive;
}

int FN(orc_linear)(const orc_cfg *c,
                   const REAL *flow, const REAL *ev, const REAL *mk, const int *n_ev,
                   const REAL *dev, const REAL *dmk, const int *n_dev,
                   REAL *loss_out, REAL *gflow, REAL *iwe_out)
{
    if (c->P > ORC_MAX_P || c->P < 1) return -1;
    const int B = c->B, H = c->H, W = c->W, P = c->P, F = c->F;
    const size_t HW = (size_t)H * W;
    orc_slot slots[ORC_MAX_SLOTS];
    int nslots = FN(linear_slots)(c, NULL);
    if (nslots > ORC_MAX_SLOTS) return -1;
    FN(linear_slots)(c, slots);
    for (int q = 0; q < nslots; ++q) if (slots[q].hi == slots[q].lo) return -2;

    FN(evset) sets[2]; int rc;
    if ((rc = FN(evset_init)(&sets[0], c, ev, mk, n_ev))) return rc;
    if ((rc = FN(evset_init)(&sets[1], c, dev, dmk, n_dev))) { FN(evset_free)(&sets[0]); return rc; }
    if (gflow) memset(gflow, 0, sizeof(REAL) * (size_t)F * P * B * 2 * HW);

    REAL *evy[2], *evx[2];
    for (int k = 0; k < 2; ++k) {
        size_t n = (size_t)(sets[k].E > 0 ? sets[k].E : 1);
        evy[k] = (REAL *)malloc(sizeof(REAL) * n); evx[k] = (REAL *)malloc(sizeof(REAL) * n);
    }
    const size_t img_sz = (size_t)B * nslots * 4 * HW;
    REAL *img = (REAL *)malloc(sizeof(REAL) * img_sz);
    ACC *aimg = (ACC *)malloc(sizeof(ACC) * img_sz), *adimg = (ACC *)malloc(sizeof(ACC) * img_sz);
    REAL *gv = gflow ? (REAL *)malloc(sizeof(REAL) * 2 * (size_t)(sets[0].E > 0 ? sets[0].E : 1)) : NULL;
    double loss_total = 0.0;

    for (int f = 0; f < F; ++f) {
        const REAL *maps = flow + (size_t)f * P * B * 2 * HW;
        /* per-event flow sampled at `update` time (loss/flow.py:266-285) */
        for (int k = 0; k < 2; ++k) {
            const FN(evset) *s = &sets[k];
            #pragma omp parallel for schedule(static)
            for (long e = 0; e < s->E; ++e) {
                const REAL *mx = maps + ((size_t)s->pass[e] * B + s->bat[e]) * 2 * HW, *my = mx + HW;
                FN(sample_flow)(mx, my, H, W, s->ev[e * 4 + 1], s->ev[e * 4 + 2], &evy[k][e], &evx[k][e], 0);
            }
        }
        #pragma omp parallel for collapse(2) schedule(dynamic, 1)
        for (int b = 0; b < B; ++b) for (int q = 0; q < nslots; ++q) {
            const orc_slot *sl = &slots[q];
            for (int k = 0; k < 2; ++k) {
                const FN(evset) *s = &sets[k];
                ACC *im = (k == 0 ? aimg : adimg) + ((size_t)b * nslots + q) * 4 * HW;
                memset(im, 0, sizeof(ACC) * 4 * HW);
                for (int corner = 0; corner < 4; ++corner)
                    for (int t = sl->lo; t < sl->hi; ++t) for (int i = 0; i < s->n[t]; ++i) {
                        long e = s->off[t] + (long)b * s->n[t] + i;
                        REAL py, px, dt;
                        REAL al = (REAL)FN(linear_pos)(c, sl, s->ts[e], s->ev[e * 4 + 1], s->ev[e * 4 + 2], evy[k][e], evx[k][e], &py, &px, &dt);
                        REAL mp = s->mk[e * 2] * al, mn = s->mk[e * 2 + 1] * al;
                        FN(corners_t) cr; FN(corners)(py, px, H, W, &cr);
                        REAL nts = (REAL)1 - R_FABS((REAL)sl->tref - s->ts[e]) / (REAL)sl->delta;
                        REAL w = cr.w[corner], wt = w * nts; long pxl = cr.pix[corner];
                        im[0 * HW + pxl] += (ACC)(REAL)(w * mp); im[1 * HW + pxl] += (ACC)(REAL)(w * mn);
                        im[2 * HW + pxl] += (ACC)(REAL)(wt * mp); im[3 * HW + pxl] += (ACC)(REAL)(wt * mn);
                    }
            }
            REAL *im = img + ((size_t)b * nslots + q) * 4 * HW;
            const ACC *ai = aimg + ((size_t)b * nslots + q) * 4 * HW, *di = adimg + ((size_t)b * nslots + q) * 4 * HW;
            for (size_t i = 0; i < 4 * HW; ++i) im[i] = (REAL)(ai[i] + di[i]);       /* loss/flow.py:725-726 */
        }
        if (iwe_out)
            memcpy(iwe_out + (size_t)f * B * nslots * 4 * HW, img, sizeof(REAL) * img_sz);

        double fl = 0.0;
        #pragma omp parallel for collapse(2) schedule(dynamic, 1) reduction(+ : fl)
        for (int b = 0; b < B; ++b) for (int q = 0; q < nslots; ++q) {
            REAL *im = img + ((size_t)b * nslots + q) * 4 * HW;
            double acc = 0.0; long cnt = 0;
            for (size_t i = 0; i < HW; ++i) {
                REAL ap = im[2 * HW + i] / (im[0 * HW + i] + (REAL)1e-9);
                REAL an = im[3 * HW + i] / (im[1 * HW + i] + (REAL)1e-9);
                acc += (double)(ap * ap) + (double)(an * an);
                cnt += ((im[0 * HW + i] + im[1 * HW + i]) != (REAL)0);
            }
            REAL den = c->loss_scaling ? (REAL)cnt + (REAL)1e-9 : (REAL)1;
            fl += acc / (double)den * slots[q].coef;
            if (gflow) {
                REAL cf = FN(upstream)(c, &slots[q], 1) / den;
                for (size_t i = 0; i < HW; ++i) for (int pol = 0; pol < 2; ++pol) {
                    REAL iw = im[pol * HW + i] + (REAL)1e-9;
                    REAL a = im[(2 + pol) * HW + i] / iw;
                    REAL ga = cf * ((REAL)2 * a);
                    im[(2 + pol) * HW + i] = ga / iw;
                    im[pol * HW + i] = -(ga * (a / iw));
                }
            }
        }
        loss_total += fl;

        if (gflow) {
            const FN(evset) *s = &sets[0]; const long E = s->E;
            REAL *gf = gflow + (size_t)f * P * B * 2 * HW;
            #pragma omp parallel for schedule(static)
            for (long e = 0; e < E; ++e) {
                const int t = s->pass[e], b = s->bat[e];
                REAL gvy = 0, gvx = 0;
                for (int q = 0; q < nslots; ++q) {
                    const orc_slot *sl = &slots[q];
                    if (t < sl->lo || t >= sl->hi) continue;
                    REAL py, px, dt;
                    REAL al = (REAL)FN(linear_pos)(c, sl, s->ts[e], s->ev[e * 4 + 1], s->ev[e * 4 + 2], evy[0][e], evx[0][e], &py, &px, &dt);
                    REAL mp = s->mk[e * 2] * al, mn = s->mk[e * 2 + 1] * al;
                    if (mp == (REAL)0 && mn == (REAL)0) continue;
                    const REAL *im = img + ((size_t)b * nslots + q) * 4 * HW;
                    REAL nts = (REAL)1 - R_FABS((REAL)sl->tref - s->ts[e]) / (REAL)sl->delta;
                    FN(corners_t) cr; FN(corners)(py, px, H, W, &cr);
                    REAL gpy = 0, gpx = 0;
                    for (int k = 0; k < 4; ++k) {
                        if (!cr.ok[k]) continue;
                        long pxl = cr.pix[k];
                        REAL gw = mp * (im[0 * HW + pxl] + nts * im[2 * HW + pxl]) + mn * (im[1 * HW + pxl] + nts * im[3 * HW + pxl]);
                        gpy += gw * FN(d1)(py, cr.cy[k]) * cr.wx[k];
                        gpx += gw * cr.wy[k] * FN(d1)(px, cr.cx[k]);
                    }
                    gvy += dt * gpy; gvx += dt * gpx;
                }
                gv[e * 2] = gvy; gv[e * 2 + 1] = gvx;
            }
            #pragma omp parallel for collapse(2) schedule(dynamic, 1)
            for (int p = 0; p < P; ++p) for (int b = 0; b < B; ++b) {
                const REAL *mx = maps + ((size_t)p * B + b) * 2 * HW, *my = mx + HW;
                REAL *gmx = gf + ((size_t)p * B + b) * 2 * HW, *gmy = gmx + HW;
                ACC *amx = (ACC *)calloc(2 * HW, sizeof(ACC)), *amy = amx + HW;
                for (int i = 0; i < s->n[p]; ++i) {
                    long e = s->off[p] + (long)b * s->n[p] + i;
                    if (gv[e * 2] == (REAL)0 && gv[e * 2 + 1] == (REAL)0) continue;
                    REAL vy, vx; FN(taps_t) tp; FN(sample_flow)(mx, my, H, W, s->ev[e * 4 + 1], s->ev[e * 4 + 2], &vy, &vx, &tp);
                    const int ty[4] = { tp.y0, tp.y0, tp.y0 + 1, tp.y0 + 1 };
                    const int tx[4] = { tp.x0, tp.x0 + 1, tp.x0, tp.x0 + 1 };
                    for (int k = 0; k < 4; ++k) if (tp.ok[k]) {
                        amy[(size_t)ty[k] * W + tx[k]] += (ACC)(REAL)(tp.w[k] * gv[e * 2]);
                        amx[(size_t)ty[k] * W + tx[k]] += (ACC)(REAL)(tp.w[k] * gv[e * 2 + 1]);
                    }
                }
                for (size_t i = 0; i < HW; ++i) { gmx[i] = (REAL)amx[i]; gmy[i] = (REAL)amy[i]; }
                free(amx);
            }
        }
    }
    *loss_out = (REAL)loss_total;
    for (int k = 0; k < 2; ++k) { free(evy[k]); free(evx[k]); }
    free(img); free(aimg); free(adimg); if (gv) free(gv);
    FN(evset_free)(&sets[0]); FN(evset_free)(&sets[1]);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Stand-alone primitives of utils/iwe.py, batched exactly like the         */
/* reference tensors.                                                       */
/* ------------------------------------------------------------------------ */

/* utils/iwe.py:17-40; mapx/mapy [B][H][W], loc [B][N][2] (y,x) -> out [B][N][2] (y,x) */
void FN(orc_get_event_flow)(const REAL *mapx, const REAL *mapy, const REAL *loc, REAL *out, int B, int N, int H, int W)
{
    const size_t HW = (size_t)H * W;
    for (int b = 0; b < B; ++b) for (int i = 0; i < N; ++i) {
        size_t e = (size_t)b * N + i;
        FN(sample_flow)(mapx + b * HW, mapy + b * HW, H, W, loc[e * 2], loc[e * 2 + 1], &out[e * 2], &out[e * 2 + 1], 0);
    }
}

/* utils/iwe.py:5-14; ts [n], loc/flow/out [n][2] */
void FN(orc_event_propagation)(const REAL *ts, const REAL *loc, const REAL *flow, REAL tref, REAL *out, long n)
{
    for (long e = 0; e < n; ++e) {
        REAL dt = tref - ts[e];
        out[e * 2] = loc[e * 2] + dt * flow[e * 2];
        out[e * 2 + 1] = loc[e * 2 + 1] + dt * flow[e * 2 + 1];
    }
}

/* utils/iwe.py:43-60; loc/mask [n][2] in place */
void FN(orc_purge_unfeasible)(REAL *loc, REAL *mask, long n, int H, int W)
{
    for (long e = 0; e < n; ++e) {
        REAL in = (REAL)FN(inside)(loc[e * 2], loc[e * 2 + 1], H, W);
        loc[e * 2] *= in; loc[e * 2 + 1] *= in; mask[e * 2] *= in; mask[e * 2 + 1] *= in;
    }
}

/* utils/iwe.py:63-113; warped [B][N][2]; bilinear: idx/w [B][4N], rounded: [B][N] */
void FN(orc_get_interpolation)(const REAL *warped, REAL *idx, REAL *w, int B, int N, int H, int W, int round_idx)
{
    for (int b = 0; b < B; ++b) for (int i = 0; i < N; ++i) {
        REAL y = warped[((size_t)b * N + i) * 2], x = warped[((size_t)b * N + i) * 2 + 1];
        if (round_idx) {
            REAL ry = R_RINT(y), rx = R_RINT(x);                       /* torch.round: half to even */
            int ok = ry >= (REAL)0 && ry < (REAL)H && rx >= (REAL)0 && rx < (REAL)W;
            idx[(size_t)b * N + i] = ok ? ry * (REAL)W + rx : (REAL)0;
            w[(size_t)b * N + i] = (REAL)ok;
        } else {
            FN(corners_t) cr; FN(corners)(y, x, H, W, &cr);
            for (int k = 0; k < 4; ++k) {
                /* idx = cy*ok*W + cx*ok as REAL arithmetic (utils/iwe.py:104,110-111) */
                REAL iy = cr.cy[k] * (REAL)cr.ok[k], ix = cr.cx[k] * (REAL)cr.ok[k];
                idx[(size_t)b * 4 * N + (size_t)k * N + i] = iy * (REAL)W + ix;
                w[(size_t)b * 4 * N + (size_t)k * N + i] = cr.wy[k] * cr.wx[k] * (REAL)cr.ok[k];
            }
        }
    }
}

/* utils/iwe.py:116-136; idx/w/pol [B][M]; iwe [B][H*W] (pre-initialised by the caller: zeros or `zeros` image) */
void FN(orc_interpolate)(const REAL *idx, const REAL *w, const REAL *pol, REAL *iwe, int B, long M, int H, int W)
{
    const size_t HW = (size_t)H * W;
    for (int b = 0; b < B; ++b) for (long i = 0; i < M; ++i) {
        REAL v = w[(size_t)b * M + i];
        if (pol) v = v * pol[(size_t)b * M + i];
        iwe[b * HW + (size_t)(long)idx[(size_t)b * M + i]] += v;
    }
}

/* utils/iwe.py:139-224 deblur_events; flow [B][2][H][W], ev [B][N][4], pol [B][N] or NULL -> iwe [B][H*W] */
void FN(orc_deblur_events)(const REAL *flow, const REAL *ev, const REAL *pol, REAL *iwe, int B, int N, int H, int W,
                           int round_idx, int round_flow)
{
    const size_t HW = (size_t)H * W;
    memset(iwe, 0, sizeof(REAL) * (size_t)B * HW);
    const int ncorner = round_idx ? 1 : 4;
    for (int b = 0; b < B; ++b) {
        const REAL *fx = flow + (size_t)b * 2 * HW, *fy = fx + HW;
        for (int corner = 0; corner < ncorner; ++corner) for (int i = 0; i < N; ++i) {
            const REAL *e = ev + ((size_t)b * N + i) * 4;
            REAL ts = e[0], y = e[1], x = e[2];
            int feas = y >= (REAL)0 && y < (REAL)H && x >= (REAL)0 && x < (REAL)W;   /* :154-160 */
            REAL qy = y * (REAL)feas, qx = x * (REAL)feas;
            REAL vy, vx;
            if (round_flow) {
                /* :185-191 : idx = y*W + x, .long() truncates */
                long id = (long)(qy * (REAL)W + qx);
                vy = fy[id]; vx = fx[id];
            } else {
                /* :164-209 manual 4-tap gather, plain multiply-adds */
                FN(corners_t) cr; FN(corners)(qy, qx, H, W, &cr);
                REAL ay = 0, ax = 0;
                for (int k = 0; k < 4; ++k) {
                    REAL iy = cr.cy[k] * (REAL)cr.ok[k], ix = cr.cx[k] * (REAL)cr.ok[k];
                    long id = (long)(iy * (REAL)W + ix);
                    REAL wk = cr.wy[k] * cr.wx[k] * (REAL)cr.ok[k];
                    if (k == 0) { ay = wk * fy[id]; ax = wk * fx[id]; }
                    else { ay = ay + wk * fy[id]; ax = ax + wk * fx[id]; }
                }
                vy = ay; vx = ax;
            }
            REAL dt = (REAL)1 - ts;
            REAL wy_ = y + dt * vy, wx_ = x + dt * vx;                 /* :214 uses the unmasked location */
            REAL pm = pol ? pol[(size_t)b * N + i] : (REAL)1;
            if (round_idx) {
                REAL ry = R_RINT(wy_), rx = R_RINT(wx_);
                int ok = ry >= (REAL)0 && ry < (REAL)H && rx >= (REAL)0 && rx < (REAL)W;
                long id = ok ? (long)(ry * (REAL)W + rx) : 0;
                iwe[b * HW + id] += ((REAL)ok * (REAL)feas) * pm;
            } else {
                FN(corners_t) cr; FN(corners)(wy_, wx_, H, W, &cr);
                iwe[b * HW + cr.pix[corner]] += (cr.w[corner] * (REAL)feas) * pm;
            }
        }
    }
}

/* index_put_ semantics of img[ys.long(), xs.long()] (dataloader/encodings.py:23-27): truncate toward zero, negative indices
   wrap once like Python indexing, anything still outside raises IndexError (return -1 here; the wrapper raises) */
static int FN(enc_pixel)(REAL xf, REAL yf, int H, int W, size_t *px)
{
    long x = (long)xf, y = (long)yf;
    if (x < 0) x += W;
    if (y < 0) y += H;
    if (x < 0 || x >= W || y < 0 || y >= H) return -1;
    *px = (size_t)y * W + (size_t)x;
    return 0;
}
/* dataloader/encodings.py:8-29; img [H][W] (+)= ps at (ys.long(), xs.long()); accumulate = 0: plain put, events in order
   (the last event of a pixel wins, like index_put_'s CPU kernel) */
int FN(orc_events_to_image)(const REAL *xs, const REAL *ys, const REAL *ps, REAL *img, long n, int H, int W, int accumulate)
{
    memset(img, 0, sizeof(REAL) * (size_t)H * W);
    for (long i = 0; i < n; ++i) {
        size_t px;
        if (FN(enc_pixel)(xs[i], ys[i], H, W, &px)) return -1;
        if (accumulate) img[px] += ps[i]; else img[px] = ps[i];
    }
    return 0;
}
/* dataloader/encodings.py:59-81; out [2][H][W] */
int FN(orc_events_to_channels)(const REAL *xs, const REAL *ys, const REAL *ps, REAL *out, long n, int H, int W)
{
    const size_t HW = (size_t)H * W;
    memset(out, 0, sizeof(REAL) * 2 * HW);
    for (long i = 0; i < n; ++i) {
        REAL p = ps[i];
        REAL mpos = p < (REAL)0 ? (REAL)0 : (p > (REAL)0 ? (REAL)1 : p);
        REAL mneg = p > (REAL)0 ? (REAL)0 : (p < (REAL)0 ? (REAL)-1 : p);
        size_t px;
        if (FN(enc_pixel)(xs[i], ys[i], H, W, &px)) return -1;
        out[px] += p * mpos; out[HW + px] += p * mneg;
    }
    return 0;
}
/* dataloader/encodings.py:32-56; out [bins][H][W] */
int FN(orc_events_to_voxel)(const REAL *xs, const REAL *ys, const REAL *ts, const REAL *ps, REAL *out, long n, int bins, int H, int W)
{
    const size_t HW = (size_t)H * W;
    memset(out, 0, sizeof(REAL) * (size_t)bins * HW);
    for (int b = 0; b < bins; ++b) for (long i = 0; i < n; ++i) {
        REAL t = ts[i] * (REAL)(bins - 1);
        REAL u = (REAL)1 - R_FABS(t - (REAL)b);
        REAL wgt = u > (REAL)0 ? u : (REAL)0;
        size_t px;
        if (FN(enc_pixel)(xs[i], ys[i], H, W, &px)) return -1;
        out[(size_t)b * HW + px] += ps[i] * wgt;
    }
    return 0;
}

/* get_hot_event_mask -- NOT in the reference: PARITY UNPINNED.  Restates the published routine of tudelft/event_flow
   (dataloader/encodings.py): sequential arg-max (first index on ties) while the rate exceeds max_rate. */
void FN(orc_get_hot_event_mask)(REAL *rate, REAL *mask, int n, int idx, int max_px, int min_obvs, REAL max_rate)
{
    for (int i = 0; i < n; ++i) mask[i] = (REAL)1;
    if (!(idx > min_obvs)) return;
    for (int r = 0; r < max_px; ++r) {
        int bi = 0;
        for (int i = 1; i < n; ++i) if (rate[i] > rate[bi]) bi = i;
        if (rate[bi] > max_rate) { rate[bi] = (REAL)0; mask[bi] = (REAL)0; } else break;
    }
}
