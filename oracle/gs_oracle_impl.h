/*
 * gs_oracle_impl.h -- body of the CPU oracle, instantiated twice by gs_oracle.c
 * (REAL=float -> suffix _f32, REAL=double -> suffix _f64).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under robosimgs_b200/ may include, link or call this.
 *
 * PARITY UNPINNED: the reference repository (Maxwell-Zhao/RoboSimGS) ships no rasterizer
 * (the only anchor is /root/reference/README.md:75, which delegates 3DGS
 * reconstruction to Nerfstudio).  The arithmetic restated here is the *published* algorithm of
 * Kerbl et al. 2023 ("3D Gaussian Splatting for Real-Time Radiance Field Rendering") with EWA
 * splatting (Zwicker et al. 2001), using the constants of the public
 * graphdeco-inria/diff-gaussian-rasterization implementation (un-vendored, un-pinned third-party
 * dependency; not present in this container).  Each function names the step of SURVEY.md section
 * 8(c) "Oracle spec" it restates.  What IS pinned by reference fixtures (camera convention) is
 * tested in tests/test_cameras.py against tests/golden/.
 *
 * Conventions (SURVEY.md 8(b) "Layouts"):
 *   view[16], proj[16]: transposed 4x4 (i.e. column-major storage of the usual matrix):
 *       x' = m[0]*x + m[4]*y + m[8]*z + m[12]
 *   quaternions (w,x,y,z) used as given (no normalisation);  shs[P][M][3];  colour out CHW.
 *   conic (A,B,C): power = -0.5*(A*dx*dx + C*dy*dy) - B*dx*dy.
 *   dL_dconic here is the TRUE derivative w.r.t. (A,B,C) (the public implementation keeps a
 *   half-weighted B internally; only final parameter gradients are contract).
 *   dL_dmean2D is in the public implementation's NDC-scaled units (pixel gradient * 0.5*W, 0.5*H).
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)
#define PARAMS CAT(GsoParams, SFX)

typedef struct {
  int P;         /* number of Gaussians */
  int sh_degree; /* active SH degree (0..3) */
  int M;         /* SH coefficients stored per channel */
  int H, W;
  REAL tanfovx, tanfovy, scale_modifier;
  REAL denom_eps; /* regulariser in the conic adjoint; 1e-7 in the public implementation */
  REAL near_plane; /* near cull on view-space z; 0.2 in the public implementation */
  REAL bg[3];
  REAL view[16];
  REAL proj[16];
  REAL campos[3];
} PARAMS;

#define TILE 16

static inline REAL FN(r_max)(REAL a, REAL b) { return a > b ? a : b; }
static inline REAL FN(r_min)(REAL a, REAL b) { return a < b ? a : b; }

/* ---- step 3: Sigma = R S S^T R^T, upper-triangular 6 (xx,xy,xz,yy,yz,zz) ---- */
static void FN(cov3d_from_scale_rot)(const REAL* s, REAL mod, const REAL* q, REAL* cov) {
  REAL r = q[0], x = q[1], y = q[2], z = q[3];
  REAL R[3][3] = {
      {1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)},
      {2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)},
      {2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)}};
  REAL d[3] = {mod * s[0], mod * s[1], mod * s[2]};
  REAL Mx[3][3];
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < 3; k++) Mx[i][k] = R[i][k] * d[k];
  REAL S[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      REAL a = 0;
      for (int k = 0; k < 3; k++) a += Mx[i][k] * Mx[j][k];
      S[i][j] = a;
    }
  cov[0] = S[0][0]; cov[1] = S[0][1]; cov[2] = S[0][2];
  cov[3] = S[1][1]; cov[4] = S[1][2]; cov[5] = S[2][2];
}

/* SH basis constants (step 9) */
#define SH_C0 ((REAL)0.28209479177387814)
#define SH_C1 ((REAL)0.4886025119029199)
static const REAL FN(SH_C2)[5] = {(REAL)1.0925484305920792, (REAL)-1.0925484305920792,
                                  (REAL)0.31539156525252005, (REAL)-1.0925484305920792,
                                  (REAL)0.5462742152960396};
static const REAL FN(SH_C3)[7] = {(REAL)-0.5900435899266435, (REAL)2.890611442640554,
                                  (REAL)-0.4570457994644658, (REAL)0.3731763325901154,
                                  (REAL)-0.4570457994644658, (REAL)1.445305721320277,
                                  (REAL)-0.5900435899266435};

/* evaluate the (deg+1)^2 basis functions at unit direction (x,y,z) */
static void FN(sh_basis)(int deg, REAL x, REAL y, REAL z, REAL* b) {
  b[0] = SH_C0;
  if (deg > 0) {
    b[1] = -SH_C1 * y; b[2] = SH_C1 * z; b[3] = -SH_C1 * x;
    if (deg > 1) {
      REAL xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      b[4] = FN(SH_C2)[0] * xy;
      b[5] = FN(SH_C2)[1] * yz;
      b[6] = FN(SH_C2)[2] * (2 * zz - xx - yy);
      b[7] = FN(SH_C2)[3] * xz;
      b[8] = FN(SH_C2)[4] * (xx - yy);
      if (deg > 2) {
        b[9] = FN(SH_C3)[0] * y * (3 * xx - yy);
        b[10] = FN(SH_C3)[1] * xy * z;
        b[11] = FN(SH_C3)[2] * y * (4 * zz - xx - yy);
        b[12] = FN(SH_C3)[3] * z * (2 * zz - 3 * xx - 3 * yy);
        b[13] = FN(SH_C3)[4] * x * (4 * zz - xx - yy);
        b[14] = FN(SH_C3)[5] * z * (xx - yy);
        b[15] = FN(SH_C3)[6] * x * (xx - 3 * yy);
      }
    }
  }
}

/* gradient of each basis function w.r.t. (x,y,z): g[k][0..2] */
static void FN(sh_basis_grad)(int deg, REAL x, REAL y, REAL z, REAL g[16][3]) {
  for (int k = 0; k < 16; k++) g[k][0] = g[k][1] = g[k][2] = 0;
  if (deg > 0) {
    g[1][1] = -SH_C1; g[2][2] = SH_C1; g[3][0] = -SH_C1;
    if (deg > 1) {
      REAL xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      const REAL* c2 = FN(SH_C2);
      g[4][0] = c2[0] * y;       g[4][1] = c2[0] * x;
      g[5][1] = c2[1] * z;       g[5][2] = c2[1] * y;
      g[6][0] = c2[2] * -2 * x;  g[6][1] = c2[2] * -2 * y;  g[6][2] = c2[2] * 4 * z;
      g[7][0] = c2[3] * z;       g[7][2] = c2[3] * x;
      g[8][0] = c2[4] * 2 * x;   g[8][1] = c2[4] * -2 * y;
      if (deg > 2) {
        const REAL* c3 = FN(SH_C3);
        g[9][0] = c3[0] * 6 * xy;             g[9][1] = c3[0] * (3 * xx - 3 * yy);
        g[10][0] = c3[1] * yz;                g[10][1] = c3[1] * xz;   g[10][2] = c3[1] * xy;
        g[11][0] = c3[2] * -2 * xy;           g[11][1] = c3[2] * (4 * zz - xx - 3 * yy);
        g[11][2] = c3[2] * 8 * yz;
        g[12][0] = c3[3] * -6 * xz;           g[12][1] = c3[3] * -6 * yz;
        g[12][2] = c3[3] * (6 * zz - 3 * xx - 3 * yy);
        g[13][0] = c3[4] * (4 * zz - 3 * xx - yy);  g[13][1] = c3[4] * -2 * xy;
        g[13][2] = c3[4] * 8 * xz;
        g[14][0] = c3[5] * 2 * xz;            g[14][1] = c3[5] * -2 * yz;
        g[14][2] = c3[5] * (xx - yy);
        g[15][0] = c3[6] * (3 * xx - 3 * yy); g[15][1] = c3[6] * -6 * xy;
      }
    }
  }
}

/* shared by forward and adjoint: view-space mean, clamped t, M = J*Wr (2x3) */
static void FN(ewa_jacobian)(const PARAMS* p, const REAL* mu, REAL* t, REAL Mjw[2][3],
                             REAL* xmask, REAL* ymask) {
  const REAL* v = p->view;
  REAL tx = v[0] * mu[0] + v[4] * mu[1] + v[8] * mu[2] + v[12];
  REAL ty = v[1] * mu[0] + v[5] * mu[1] + v[9] * mu[2] + v[13];
  REAL tz = v[2] * mu[0] + v[6] * mu[1] + v[10] * mu[2] + v[14];
  REAL limx = (REAL)1.3 * p->tanfovx, limy = (REAL)1.3 * p->tanfovy;
  REAL txtz = tx / tz, tytz = ty / tz;
  *xmask = (txtz < -limx || txtz > limx) ? 0 : 1;
  *ymask = (tytz < -limy || tytz > limy) ? 0 : 1;
  tx = FN(r_min)(limx, FN(r_max)(-limx, txtz)) * tz;
  ty = FN(r_min)(limy, FN(r_max)(-limy, tytz)) * tz;
  t[0] = tx; t[1] = ty; t[2] = tz;
  REAL fx = p->W / (2 * p->tanfovx), fy = p->H / (2 * p->tanfovy);
  REAL J[2][3] = {{fx / tz, 0, -(fx * tx) / (tz * tz)}, {0, fy / tz, -(fy * ty) / (tz * tz)}};
  /* Wr[i][k] = view[4k+i] (rotation part, row i) */
  for (int i = 0; i < 2; i++)
    for (int k = 0; k < 3; k++)
      Mjw[i][k] = J[i][0] * v[4 * k + 0] + J[i][1] * v[4 * k + 1] + J[i][2] * v[4 * k + 2];
}

/*
 * Steps 1-9 of the oracle spec: cull, Sigma3D, EWA Sigma2D (+0.3), conic, radius, pixel centre,
 * tile rect, SH colour.  One Gaussian per loop iteration.
 */
int FN(gso_preprocess)(const PARAMS* p, const REAL* means, const REAL* shs,
                       const REAL* colors_precomp, const REAL* opac, const REAL* scales,
                       const REAL* rots, const REAL* cov3d_precomp, int* radii, REAL* xy,
                       REAL* depths, REAL* cov3d, REAL* rgb, REAL* conic_opacity,
                       int* tiles_touched, unsigned char* clamped) {
  const int P = p->P;
  const int gx = (p->W + TILE - 1) / TILE, gy = (p->H + TILE - 1) / TILE;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    radii[i] = 0;
    tiles_touched[i] = 0;
    const REAL* mu = means + 3 * i;
    const REAL* v = p->view;
    const REAL* pm = p->proj;
    /* step 1: near cull on view-space z */
    REAL vz = v[2] * mu[0] + v[6] * mu[1] + v[10] * mu[2] + v[14];
    if (vz <= p->near_plane) continue;
    /* step 2 */
    REAL hx = pm[0] * mu[0] + pm[4] * mu[1] + pm[8] * mu[2] + pm[12];
    REAL hy = pm[1] * mu[0] + pm[5] * mu[1] + pm[9] * mu[2] + pm[13];
    REAL hw = pm[3] * mu[0] + pm[7] * mu[1] + pm[11] * mu[2] + pm[15];
    REAL pw = 1 / (hw + (REAL)0.0000001);
    REAL ndcx = hx * pw, ndcy = hy * pw;
    /* step 3 */
    REAL* c3 = cov3d + 6 * i;
    if (cov3d_precomp) {
      for (int k = 0; k < 6; k++) c3[k] = cov3d_precomp[6 * i + k];
    } else {
      FN(cov3d_from_scale_rot)(scales + 3 * i, p->scale_modifier, rots + 4 * i, c3);
    }
    /* step 4 */
    REAL t[3], Mjw[2][3], xm, ym;
    FN(ewa_jacobian)(p, mu, t, Mjw, &xm, &ym);
    REAL S[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
    REAL MS[2][3];
    for (int a = 0; a < 2; a++)
      for (int k = 0; k < 3; k++)
        MS[a][k] = Mjw[a][0] * S[0][k] + Mjw[a][1] * S[1][k] + Mjw[a][2] * S[2][k];
    REAL ca = MS[0][0] * Mjw[0][0] + MS[0][1] * Mjw[0][1] + MS[0][2] * Mjw[0][2] + (REAL)0.3;
    REAL cb = MS[0][0] * Mjw[1][0] + MS[0][1] * Mjw[1][1] + MS[0][2] * Mjw[1][2];
    REAL cc = MS[1][0] * Mjw[1][0] + MS[1][1] * Mjw[1][1] + MS[1][2] * Mjw[1][2] + (REAL)0.3;
    /* step 5 */
    REAL det = ca * cc - cb * cb;
    if (det == 0) continue;
    REAL det_inv = 1 / det;
    REAL A = cc * det_inv, B = -cb * det_inv, C = ca * det_inv;
    /* step 6 */
    REAL mid = (REAL)0.5 * (ca + cc);
    REAL disc = SQRT(FN(r_max)((REAL)0.1, mid * mid - det));
    REAL l1 = mid + disc, l2 = mid - disc;
    REAL rad = CEIL(3 * SQRT(FN(r_max)(l1, l2)));
    /* step 7 */
    REAL px = ((ndcx + 1) * p->W - 1) * (REAL)0.5;
    REAL py = ((ndcy + 1) * p->H - 1) * (REAL)0.5;
    /* step 8 */
    int irad = (int)rad;
    int x0 = (int)((px - irad) / TILE), y0 = (int)((py - irad) / TILE);
    int x1 = (int)((px + irad + TILE - 1) / TILE), y1 = (int)((py + irad + TILE - 1) / TILE);
    x0 = x0 < 0 ? 0 : (x0 > gx ? gx : x0);
    y0 = y0 < 0 ? 0 : (y0 > gy ? gy : y0);
    x1 = x1 < 0 ? 0 : (x1 > gx ? gx : x1);
    y1 = y1 < 0 ? 0 : (y1 > gy ? gy : y1);
    if ((x1 - x0) * (y1 - y0) == 0) continue;
    /* step 9 */
    if (colors_precomp) {
      for (int c = 0; c < 3; c++) rgb[3 * i + c] = colors_precomp[3 * i + c];
    } else {
      REAL dx = mu[0] - p->campos[0], dy = mu[1] - p->campos[1], dz = mu[2] - p->campos[2];
      REAL inv = 1 / SQRT(dx * dx + dy * dy + dz * dz);
      REAL b[16];
      FN(sh_basis)(p->sh_degree, dx * inv, dy * inv, dz * inv, b);
      int nb = (p->sh_degree + 1) * (p->sh_degree + 1);
      const REAL* sh = shs + (size_t)i * p->M * 3;
      for (int c = 0; c < 3; c++) {
        REAL acc = 0;
        for (int k = 0; k < nb; k++) acc += b[k] * sh[3 * k + c];
        acc += (REAL)0.5;
        clamped[3 * i + c] = acc < 0;
        rgb[3 * i + c] = acc < 0 ? 0 : acc;
      }
    }
    depths[i] = vz;
    radii[i] = irad;
    xy[2 * i] = px; xy[2 * i + 1] = py;
    conic_opacity[4 * i] = A; conic_opacity[4 * i + 1] = B; conic_opacity[4 * i + 2] = C;
    conic_opacity[4 * i + 3] = opac[i];
    tiles_touched[i] = (x1 - x0) * (y1 - y0);
  }
  return 0;
}

/* view-space z > 0.2 (markVisible / checkFrustum) */
int FN(gso_mark_visible)(const PARAMS* p, const REAL* means, unsigned char* present) {
  const REAL* v = p->view;
  for (int i = 0; i < p->P; i++) {
    const REAL* mu = means + 3 * i;
    REAL vz = v[2] * mu[0] + v[6] * mu[1] + v[10] * mu[2] + v[14];
    present[i] = vz > (REAL)0.2;
  }
  return 0;
}

typedef struct {
  unsigned int tile;
  int idx;
  long seq;
  REAL depth;
} FN(PairRec);

static int FN(pair_cmp)(const void* a, const void* b) {
  const FN(PairRec)* x = (const FN(PairRec)*)a;
  const FN(PairRec)* y = (const FN(PairRec)*)b;
  if (x->tile != y->tile) return x->tile < y->tile ? -1 : 1;
  if (x->depth != y->depth) return x->depth < y->depth ? -1 : 1;
  return x->seq < y->seq ? -1 : (x->seq > y->seq ? 1 : 0);
}

/*
 * Step 10: duplicate every Gaussian into the tiles of its rect, stable-sort by (tile, depth),
 * emit per-tile [start,end).  f32: LSD radix on (tile<<32 | depth bits) exactly like the public
 * implementation; f64: comparison sort with emit order as tie-break (== stable).
 * Returns number of pairs written (must equal sum(tiles_touched)).
 */
long FN(gso_bin)(const PARAMS* p, const int* radii, const REAL* xy, const REAL* depths,
                 long D, int* point_list, int* ranges) {
  const int gx = (p->W + TILE - 1) / TILE, gy = (p->H + TILE - 1) / TILE;
  const int T = gx * gy;
  for (int t = 0; t < 2 * T; t++) ranges[t] = 0;
  if (D == 0) return 0;
  FN(PairRec)* recs = (FN(PairRec)*)malloc(sizeof(FN(PairRec)) * (size_t)D);
  long n = 0;
  for (int i = 0; i < p->P; i++) {
    if (radii[i] <= 0) continue;
    REAL px = xy[2 * i], py = xy[2 * i + 1];
    int irad = radii[i];
    int x0 = (int)((px - irad) / TILE), y0 = (int)((py - irad) / TILE);
    int x1 = (int)((px + irad + TILE - 1) / TILE), y1 = (int)((py + irad + TILE - 1) / TILE);
    x0 = x0 < 0 ? 0 : (x0 > gx ? gx : x0);
    y0 = y0 < 0 ? 0 : (y0 > gy ? gy : y0);
    x1 = x1 < 0 ? 0 : (x1 > gx ? gx : x1);
    y1 = y1 < 0 ? 0 : (y1 > gy ? gy : y1);
    for (int y = y0; y < y1; y++)
      for (int x = x0; x < x1; x++) {
        if (n >= D) { free(recs); return -1; }
        recs[n].tile = (unsigned)(y * gx + x);
        recs[n].idx = i;
        recs[n].seq = n;
        recs[n].depth = depths[i];
        n++;
      }
  }
#if REAL_IS_FLOAT
  {
    /* LSD radix sort, 8-bit digits, on 64-bit keys; stable */
    unsigned long long* k0 = (unsigned long long*)malloc(8 * (size_t)n);
    unsigned long long* k1 = (unsigned long long*)malloc(8 * (size_t)n);
    int* v0 = (int*)malloc(4 * (size_t)n);
    int* v1 = (int*)malloc(4 * (size_t)n);
    for (long j = 0; j < n; j++) {
      unsigned int db;
      float d = recs[j].depth;
      memcpy(&db, &d, 4);
      k0[j] = ((unsigned long long)recs[j].tile << 32) | db;
      v0[j] = recs[j].idx;
    }
    int tile_bits = 0;
    while ((1 << tile_bits) < T) tile_bits++;
    int nbits = 32 + tile_bits;
    for (int shift = 0; shift < nbits; shift += 8) {
      size_t cnt[257];
      memset(cnt, 0, sizeof(cnt));
      for (long j = 0; j < n; j++) cnt[((k0[j] >> shift) & 255) + 1]++;
      for (int b = 0; b < 256; b++) cnt[b + 1] += cnt[b];
      for (long j = 0; j < n; j++) {
        size_t d = cnt[(k0[j] >> shift) & 255]++;
        k1[d] = k0[j];
        v1[d] = v0[j];
      }
      unsigned long long* tk = k0; k0 = k1; k1 = tk;
      int* tv = v0; v0 = v1; v1 = tv;
    }
    for (long j = 0; j < n; j++) {
      point_list[j] = v0[j];
      recs[j].tile = (unsigned)(k0[j] >> 32);
    }
    free(k0); free(k1); free(v0); free(v1);
  }
#else
  qsort(recs, (size_t)n, sizeof(FN(PairRec)), FN(pair_cmp));
  for (long j = 0; j < n; j++) point_list[j] = recs[j].idx;
#endif
  for (long j = 0; j < n; j++) {
    unsigned t = recs[j].tile;
    if (j == 0 || recs[j - 1].tile != t) ranges[2 * t] = (int)j;
    if (j == n - 1 || recs[j + 1].tile != t) ranges[2 * t + 1] = (int)(j + 1);
  }
  free(recs);
  return n;
}

/* Step 11: per-pixel front-to-back compositing. */
int FN(gso_render)(const PARAMS* p, const int* ranges, const int* point_list, const REAL* xy,
                   const REAL* rgb, const REAL* conic_opacity, REAL* out_color, REAL* final_T,
                   int* n_contrib) {
  const int W = p->W, H = p->H;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(dynamic, 1)
  for (int tile = 0; tile < gx * gy; tile++) {
    int tx = tile % gx, ty = tile / gx;
    int r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
    for (int ly = 0; ly < TILE; ly++)
      for (int lx = 0; lx < TILE; lx++) {
        int ix = tx * TILE + lx, iy = ty * TILE + ly;
        if (ix >= W || iy >= H) continue;
        REAL pxf = (REAL)ix, pyf = (REAL)iy;
        REAL T = 1, C[3] = {0, 0, 0};
        int contributor = 0, last = 0;
        for (int j = r0; j < r1; j++) {
          contributor++;
          int g = point_list[j];
          REAL dx = xy[2 * g] - pxf, dy = xy[2 * g + 1] - pyf;
          const REAL* co = conic_opacity + 4 * g;
          REAL power = (REAL)-0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
          if (power > 0) continue;
          REAL alpha = FN(r_min)((REAL)0.99, co[3] * EXP(power));
          if (alpha < (REAL)1.0 / 255) continue;
          REAL test_T = T * (1 - alpha);
          if (test_T < (REAL)0.0001) break;
          for (int c = 0; c < 3; c++) C[c] += rgb[3 * g + c] * alpha * T;
          T = test_T;
          last = contributor;
        }
        size_t pid = (size_t)iy * W + ix;
        final_T[pid] = T;
        n_contrib[pid] = last;
        for (int c = 0; c < 3; c++) out_color[(size_t)c * H * W + pid] = C[c] + T * p->bg[c];
      }
  }
  return 0;
}

static inline void FN(atomic_add)(REAL* dst, REAL v) {
#pragma omp atomic
  *dst += v;
}

/*
 * Adjoint of step 11: back-to-front replay per pixel.  Outputs must be
 * zero-initialised by the caller: dL_dmean2D[P][2] (NDC-scaled), dL_dconic[P][3] (true d/dA,B,C),
 * dL_dopacity[P], dL_dcolor[P][3].
 */
int FN(gso_render_backward)(const PARAMS* p, const int* ranges, const int* point_list,
                            const REAL* xy, const REAL* conic_opacity, const REAL* rgb,
                            const REAL* final_T, const int* n_contrib, const REAL* dL_dpix,
                            REAL* dL_dmean2D, REAL* dL_dconic, REAL* dL_dopacity,
                            REAL* dL_dcolor) {
  const int W = p->W, H = p->H;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const REAL ddelx_dx = (REAL)0.5 * W, ddely_dy = (REAL)0.5 * H;
#pragma omp parallel for schedule(dynamic, 1)
  for (int tile = 0; tile < gx * gy; tile++) {
    int tx = tile % gx, ty = tile / gx;
    int r0 = ranges[2 * tile];
    for (int ly = 0; ly < TILE; ly++)
      for (int lx = 0; lx < TILE; lx++) {
        int ix = tx * TILE + lx, iy = ty * TILE + ly;
        if (ix >= W || iy >= H) continue;
        size_t pid = (size_t)iy * W + ix;
        REAL pxf = (REAL)ix, pyf = (REAL)iy;
        const REAL T_final = final_T[pid];
        REAL T = T_final;
        REAL dpix[3], bg_dot = 0;
        for (int c = 0; c < 3; c++) {
          dpix[c] = dL_dpix[(size_t)c * H * W + pid];
          bg_dot += p->bg[c] * dpix[c];
        }
        REAL accum[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0;
        for (int k = n_contrib[pid] - 1; k >= 0; k--) {
          int g = point_list[r0 + k];
          REAL dx = xy[2 * g] - pxf, dy = xy[2 * g + 1] - pyf;
          const REAL* co = conic_opacity + 4 * g;
          REAL power = (REAL)-0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
          if (power > 0) continue;
          REAL G = EXP(power);
          REAL alpha = FN(r_min)((REAL)0.99, co[3] * G);
          if (alpha < (REAL)1.0 / 255) continue;
          T = T / (1 - alpha);
          REAL w = alpha * T;
          REAL dL_dalpha = 0;
          for (int c = 0; c < 3; c++) {
            REAL col = rgb[3 * g + c];
            accum[c] = last_alpha * last_color[c] + (1 - last_alpha) * accum[c];
            last_color[c] = col;
            dL_dalpha += (col - accum[c]) * dpix[c];
            FN(atomic_add)(dL_dcolor + 3 * g + c, w * dpix[c]);
          }
          dL_dalpha *= T;
          last_alpha = alpha;
          dL_dalpha += (-T_final / (1 - alpha)) * bg_dot;
          REAL dL_dG = co[3] * dL_dalpha;
          REAL gdx = G * dx, gdy = G * dy;
          REAL dG_ddx = -gdx * co[0] - gdy * co[1];
          REAL dG_ddy = -gdy * co[2] - gdx * co[1];
          FN(atomic_add)(dL_dmean2D + 2 * g, dL_dG * dG_ddx * ddelx_dx);
          FN(atomic_add)(dL_dmean2D + 2 * g + 1, dL_dG * dG_ddy * ddely_dy);
          FN(atomic_add)(dL_dconic + 3 * g, (REAL)-0.5 * gdx * dx * dL_dG);
          FN(atomic_add)(dL_dconic + 3 * g + 1, -gdx * dy * dL_dG);
          FN(atomic_add)(dL_dconic + 3 * g + 2, (REAL)-0.5 * gdy * dy * dL_dG);
          FN(atomic_add)(dL_dopacity + g, G * dL_dalpha);
        }
      }
  }
  return 0;
}

/*
 * Adjoint of steps 9..2, one Gaussian per iteration.  Gaussians with
 * radii <= 0 receive zero gradients.  Every output is overwritten.
 *   dL_dcov3D[P][6] is the derivative w.r.t. the 6 unique entries (off-diagonals counted once
 *   each, i.e. they carry the sum over both symmetric positions).
 */
int FN(gso_preprocess_backward)(const PARAMS* p, const REAL* means, const REAL* shs,
                                const REAL* scales, const REAL* rots, const REAL* cov3d,
                                const int* radii, const unsigned char* clamped,
                                const REAL* dL_dmean2D, const REAL* dL_dconic,
                                const REAL* dL_dcolor, REAL* dL_dmeans, REAL* dL_dshs,
                                REAL* dL_dscales, REAL* dL_drots, REAL* dL_dcov3D) {
  const int P = p->P;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    REAL* gm = dL_dmeans + 3 * i;
    gm[0] = gm[1] = gm[2] = 0;
    REAL* gcov = dL_dcov3D + 6 * i;
    for (int k = 0; k < 6; k++) gcov[k] = 0;
    if (dL_dshs)
      for (int k = 0; k < p->M * 3; k++) dL_dshs[(size_t)i * p->M * 3 + k] = 0;
    if (dL_dscales) {
      for (int k = 0; k < 3; k++) dL_dscales[3 * i + k] = 0;
      for (int k = 0; k < 4; k++) dL_drots[4 * i + k] = 0;
    }
    if (radii[i] <= 0) continue;
    const REAL* mu = means + 3 * i;

    /* ---- A.2: conic -> Sigma2D -> Sigma3D and view-space mean ---- */
    REAL t[3], Mjw[2][3], xm, ym;
    FN(ewa_jacobian)(p, mu, t, Mjw, &xm, &ym);
    const REAL* c3 = cov3d + 6 * i;
    REAL S[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
    REAL MS[2][3];
    for (int a = 0; a < 2; a++)
      for (int k = 0; k < 3; k++)
        MS[a][k] = Mjw[a][0] * S[0][k] + Mjw[a][1] * S[1][k] + Mjw[a][2] * S[2][k];
    REAL a_ = MS[0][0] * Mjw[0][0] + MS[0][1] * Mjw[0][1] + MS[0][2] * Mjw[0][2] + (REAL)0.3;
    REAL b_ = MS[0][0] * Mjw[1][0] + MS[0][1] * Mjw[1][1] + MS[0][2] * Mjw[1][2];
    REAL c_ = MS[1][0] * Mjw[1][0] + MS[1][1] * Mjw[1][1] + MS[1][2] * Mjw[1][2] + (REAL)0.3;
    REAL det = a_ * c_ - b_ * b_;
    REAL d2inv = 1 / (det * det + p->denom_eps);
    REAL gA = dL_dconic[3 * i], gB = dL_dconic[3 * i + 1], gC = dL_dconic[3 * i + 2];
    /* (A,B,C) = (c,-b,a)/det.  Derivatives w.r.t. the unique entries a,b,c of Sigma2D: */
    REAL ga = d2inv * (-c_ * c_ * gA + b_ * c_ * gB + (det - a_ * c_) * gC);
    REAL gc = d2inv * (-a_ * a_ * gC + a_ * b_ * gB + (det - a_ * c_) * gA);
    REAL gb = d2inv * (2 * b_ * c_ * gA - (det + 2 * b_ * b_) * gB + 2 * a_ * b_ * gC);
    /* symmetric 2x2 gradient matrix G2 = [[ga, gb/2],[gb/2, gc]] */
    REAL G2[2][2] = {{ga, (REAL)0.5 * gb}, {(REAL)0.5 * gb, gc}};
    /* dL/dSigma3D = M^T G2 M (full symmetric); unique-entry derivative doubles off-diagonals */
    REAL G2M[2][3];
    for (int a = 0; a < 2; a++)
      for (int k = 0; k < 3; k++) G2M[a][k] = G2[a][0] * Mjw[0][k] + G2[a][1] * Mjw[1][k];
    REAL GS[3][3];
    for (int r = 0; r < 3; r++)
      for (int k = 0; k < 3; k++) GS[r][k] = Mjw[0][r] * G2M[0][k] + Mjw[1][r] * G2M[1][k];
    gcov[0] = GS[0][0]; gcov[3] = GS[1][1]; gcov[5] = GS[2][2];
    gcov[1] = 2 * GS[0][1]; gcov[2] = 2 * GS[0][2]; gcov[4] = 2 * GS[1][2];
    /* dL/dM = 2 G2 M Sigma */
    REAL gM[2][3];
    for (int a = 0; a < 2; a++)
      for (int k = 0; k < 3; k++)
        gM[a][k] = 2 * (G2M[a][0] * S[0][k] + G2M[a][1] * S[1][k] + G2M[a][2] * S[2][k]);
    /* M = J Wr  =>  dL/dJ[a][j] = sum_k gM[a][k] * Wr[j][k],  Wr[j][k] = view[4k+j] */
    const REAL* v = p->view;
    REAL gJ[2][3];
    for (int a = 0; a < 2; a++)
      for (int j = 0; j < 3; j++)
        gJ[a][j] = gM[a][0] * v[j] + gM[a][1] * v[4 + j] + gM[a][2] * v[8 + j];
    REAL fx = p->W / (2 * p->tanfovx), fy = p->H / (2 * p->tanfovy);
    REAL tz = 1 / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
    REAL gtx = xm * (-fx * tz2 * gJ[0][2]);
    REAL gty = ym * (-fy * tz2 * gJ[1][2]);
    REAL gtz = -fx * tz2 * gJ[0][0] - fy * tz2 * gJ[1][1] + (2 * fx * t[0]) * tz3 * gJ[0][2] +
               (2 * fy * t[1]) * tz3 * gJ[1][2];
    /* t = Wr mu + trans  =>  dL/dmu = Wr^T gt */
    for (int k = 0; k < 3; k++) gm[k] += v[4 * k] * gtx + v[4 * k + 1] * gty + v[4 * k + 2] * gtz;

    /* ---- A.3: pixel centre through the full projection ---- */
    const REAL* pm = p->proj;
    REAL hx = pm[0] * mu[0] + pm[4] * mu[1] + pm[8] * mu[2] + pm[12];
    REAL hy = pm[1] * mu[0] + pm[5] * mu[1] + pm[9] * mu[2] + pm[13];
    REAL hw = pm[3] * mu[0] + pm[7] * mu[1] + pm[11] * mu[2] + pm[15];
    REAL w = 1 / (hw + (REAL)0.0000001);
    REAL g2x = dL_dmean2D[2 * i], g2y = dL_dmean2D[2 * i + 1];
    for (int k = 0; k < 3; k++) {
      REAL dnx = pm[4 * k] * w - pm[4 * k + 3] * hx * w * w;
      REAL dny = pm[4 * k + 1] * w - pm[4 * k + 3] * hy * w * w;
      gm[k] += dnx * g2x + dny * g2y;
    }

    /* ---- A.4: SH colour ---- */
    if (shs && dL_dshs) {
      REAL dir[3] = {mu[0] - p->campos[0], mu[1] - p->campos[1], mu[2] - p->campos[2]};
      REAL len = SQRT(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
      REAL inv = 1 / len;
      REAL x = dir[0] * inv, y = dir[1] * inv, z = dir[2] * inv;
      REAL b[16], bg_[16][3];
      FN(sh_basis)(p->sh_degree, x, y, z, b);
      FN(sh_basis_grad)(p->sh_degree, x, y, z, bg_);
      int nb = (p->sh_degree + 1) * (p->sh_degree + 1);
      const REAL* sh = shs + (size_t)i * p->M * 3;
      REAL* gsh = dL_dshs + (size_t)i * p->M * 3;
      REAL gd[3] = {0, 0, 0};
      for (int c = 0; c < 3; c++) {
        REAL gcol = clamped[3 * i + c] ? 0 : dL_dcolor[3 * i + c];
        for (int k = 0; k < nb; k++) {
          gsh[3 * k + c] = b[k] * gcol;
          REAL s = sh[3 * k + c] * gcol;
          gd[0] += bg_[k][0] * s; gd[1] += bg_[k][1] * s; gd[2] += bg_[k][2] * s;
        }
      }
      REAL dot = x * gd[0] + y * gd[1] + z * gd[2];
      gm[0] += (gd[0] - x * dot) * inv;
      gm[1] += (gd[1] - y * dot) * inv;
      gm[2] += (gd[2] - z * dot) * inv;
    }

    /* ---- A.5: Sigma3D -> scale, quaternion ---- */
    if (scales && dL_dscales) {
      const REAL* q = rots + 4 * i;
      REAL r = q[0], qx = q[1], qy = q[2], qz = q[3];
      REAL R[3][3] = {
          {1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - r * qz), 2 * (qx * qz + r * qy)},
          {2 * (qx * qy + r * qz), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - r * qx)},
          {2 * (qx * qz - r * qy), 2 * (qy * qz + r * qx), 1 - 2 * (qx * qx + qy * qy)}};
      REAL mod = p->scale_modifier;
      REAL d[3] = {mod * scales[3 * i], mod * scales[3 * i + 1], mod * scales[3 * i + 2]};
      /* full symmetric dL/dSigma */
      REAL Gf[3][3] = {{gcov[0], (REAL)0.5 * gcov[1], (REAL)0.5 * gcov[2]},
                       {(REAL)0.5 * gcov[1], gcov[3], (REAL)0.5 * gcov[4]},
                       {(REAL)0.5 * gcov[2], (REAL)0.5 * gcov[4], gcov[5]}};
      /* Mx = R diag(d); Sigma = Mx Mx^T; dL/dMx = 2 Gf Mx */
      REAL gMx[3][3];
      for (int a = 0; a < 3; a++)
        for (int k = 0; k < 3; k++)
          gMx[a][k] = 2 * (Gf[a][0] * R[0][k] + Gf[a][1] * R[1][k] + Gf[a][2] * R[2][k]) * d[k];
      REAL F[3][3];
      for (int k = 0; k < 3; k++) {
        REAL acc = 0;
        for (int a = 0; a < 3; a++) {
          acc += gMx[a][k] * R[a][k];
          F[a][k] = gMx[a][k] * d[k];
        }
        dL_dscales[3 * i + k] = mod * acc;
      }
      REAL* gq = dL_drots + 4 * i;
      gq[0] = 2 * (-qz * F[0][1] + qy * F[0][2] + qz * F[1][0] - qx * F[1][2] - qy * F[2][0] +
                   qx * F[2][1]);
      gq[1] = 2 * (qy * F[0][1] + qz * F[0][2] + qy * F[1][0] - 2 * qx * F[1][1] - r * F[1][2] +
                   qz * F[2][0] + r * F[2][1] - 2 * qx * F[2][2]);
      gq[2] = 2 * (-2 * qy * F[0][0] + qx * F[0][1] + r * F[0][2] + qx * F[1][0] + qz * F[1][2] -
                   r * F[2][0] + qz * F[2][1] - 2 * qy * F[2][2]);
      gq[3] = 2 * (-2 * qz * F[0][0] - r * F[0][1] + qx * F[0][2] + r * F[1][0] -
                   2 * qz * F[1][1] + qy * F[1][2] + qx * F[2][0] + qy * F[2][1]);
    }
  }
  return 0;
}

#undef PARAMS
#undef FN
#undef CAT
#undef CAT_
