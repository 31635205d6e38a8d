// render_sim.c -- CPU model of the compositing kernel's work (records walked, cull batches, block
// survivors, pixel contributions) for alternative block shapes.  Analysis tool only (tools/), uses
// projected splats produced by the oracle; never part of the product path.
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float depth; int id; } Ent;
static int cmp_ent(const void* a, const void* b) {
  const Ent *x = a, *y = b;
  if (x->depth < y->depth) return -1;
  if (x->depth > y->depth) return 1;
  return x->id - y->id;
}
static inline float clampf(float v, float lo, float hi) { return fminf(hi, fmaxf(lo, v)); }
static int block_may(float x, float y, float A, float B, float C, float thr, float x0, float y0, float x1, float y1) {
  float cx = clampf(x, x0, x1), cy = clampf(y, y0, y1);
  float dxe = cx - x, dye = cy - y;
  float dy1 = clampf(y - B * dxe / C, y0, y1) - y;
  float q1 = A * dxe * dxe + 2.f * B * dxe * dy1 + C * dy1 * dy1;
  float dx2 = clampf(x - B * dye / A, x0, x1) - x;
  float q2 = A * dx2 * dx2 + 2.f * B * dx2 * dye + C * dye * dye;
  float q = fminf(q1, q2);
  return 0.5f * q <= thr * 1.00001f + 1e-5f;
}
static int in_ref_rect(float px, float py, float fr, int tx, int ty, int gx, int gy) {
  int x0 = (int)((px - fr) / 16); if (x0 < 0) x0 = 0; if (x0 > gx) x0 = gx;
  int y0 = (int)((py - fr) / 16); if (y0 < 0) y0 = 0; if (y0 > gy) y0 = gy;
  int x1 = (int)((px + fr + 15) / 16); if (x1 < 0) x1 = 0; if (x1 > gx) x1 = gx;
  int y1 = (int)((py + fr + 15) / 16); if (y1 < 0) y1 = 0; if (y1 > gy) y1 = gy;
  return tx >= x0 && tx < x1 && ty >= y0 && ty < y1;
}

#define NSHAPE 6
static const int SH_W[NSHAPE] = {8, 16, 8, 16, 16, 16};
static const int SH_H[NSHAPE] = {4, 2, 8, 4, 8, 16};

// out: [0]=pairs, [1]=chunks walked, [2]=tile survivors, [3]=pixel contributions, [4]=records walked (exact stop),
//      [5]=pixel evaluations that pass power/alpha but pixel already done(unused),
//      [8+3*s+0]=cull batches, [8+3*s+1]=block survivors evaluated, [8+3*s+2]=lanes useful in those evals
int render_sim(int P, const float* xy, const float* co, const int* radii, const float* depths, int W, int H,
               int bin_shift, int mode_bwd, double* out) {
  const int gx = (W + 15) / 16, gy = (H + 15) / 16;
  const int bt = 16 << bin_shift;
  const int gbx = (W + bt - 1) / bt, gby = (H + bt - 1) / bt;
  const int nb = gbx * gby;
  int* cnt = calloc(nb + 1, sizeof(int));
  float* thr = malloc(sizeof(float) * P);
  // pass 1: count
  for (int pass = 0; pass < 2; pass++) {
    static Ent* ents; static int* start; static int* fill;
    if (pass == 1) {
      start = malloc(sizeof(int) * (nb + 1));
      start[0] = 0;
      for (int b = 0; b < nb; b++) start[b + 1] = start[b] + cnt[b];
      ents = malloc(sizeof(Ent) * (size_t)(start[nb] + 1));
      fill = calloc(nb, sizeof(int));
      out[0] = start[nb];
    }
    for (int i = 0; i < P; i++) {
      if (radii[i] <= 0) continue;
      const float A = co[4 * i], B = co[4 * i + 1], C = co[4 * i + 2], o = co[4 * i + 3];
      if (!(o > 0)) continue;
      const float t = logf(255.f * o) + 0.01f;
      thr[i] = t;
      if (!(t > 0)) continue;
      const float det = A * C - B * B;
      if (!(det > 0)) continue;
      const float x = xy[2 * i], y = xy[2 * i + 1], fr = (float)radii[i];
      const float xe = sqrtf(2 * t * C / det) * 1.005f + 0.02f, ye = sqrtf(2 * t * A / det) * 1.005f + 0.02f;
      int tx0 = (int)((x - fr) / 16), ty0 = (int)((y - fr) / 16), tx1 = (int)((x + fr + 15) / 16), ty1 = (int)((y + fr + 15) / 16);
      if (tx0 < 0) tx0 = 0; if (ty0 < 0) ty0 = 0; if (tx1 > gx) tx1 = gx; if (ty1 > gy) ty1 = gy;
      if (tx0 > gx) tx0 = gx; if (ty0 > gy) ty0 = gy; if (tx1 < 0) tx1 = 0; if (ty1 < 0) ty1 = 0;
      if (tx1 <= tx0 || ty1 <= ty0) continue;
      // ellipse AABB in pixels intersected with the reference rect (in pixels)
      float X0 = fmaxf(x - xe, tx0 * 16.f), X1 = fminf(x + xe, tx1 * 16.f - 1.f);
      float Y0 = fmaxf(y - ye, ty0 * 16.f), Y1 = fminf(y + ye, ty1 * 16.f - 1.f);
      if (X1 < X0 || Y1 < Y0) continue;
      int bx0 = (int)floorf(X0 / bt), bx1 = (int)floorf(X1 / bt), by0 = (int)floorf(Y0 / bt), by1 = (int)floorf(Y1 / bt);
      if (bx0 < 0) bx0 = 0; if (by0 < 0) by0 = 0; if (bx1 >= gbx) bx1 = gbx - 1; if (by1 >= gby) by1 = gby - 1;
      for (int by = by0; by <= by1; by++)
        for (int bx = bx0; bx <= bx1; bx++) {
          // exact-ish: ellipse vs bin block
          if (!block_may(x, y, A, B, C, t, bx * (float)bt, by * (float)bt, bx * (float)bt + bt - 1, by * (float)bt + bt - 1)) continue;
          const int b = by * gbx + bx;
          if (pass == 0) cnt[b]++;
          else { Ent e = {depths[i], i}; ents[start[b] + fill[b]++] = e; }
        }
    }
    if (pass == 1) {
#pragma omp parallel for schedule(dynamic, 1)
      for (int b = 0; b < nb; b++) qsort(ents + start[b], cnt[b], sizeof(Ent), cmp_ent);
      double acc[64];
      memset(acc, 0, sizeof(acc));
#pragma omp parallel
      {
        double loc[64];
        memset(loc, 0, sizeof(loc));
#pragma omp for schedule(dynamic, 4)
        for (int tile = 0; tile < gx * gy; tile++) {
          const int tx = tile % gx, ty = tile / gx;
          const int b = (ty >> bin_shift) * gbx + (tx >> bin_shift);
          const Ent* L = ents + start[b];
          const int n = cnt[b];
          float T[256]; unsigned char done[256];
          int ndone = 0;
          for (int p = 0; p < 256; p++) {
            const int px = tx * 16 + (p & 15), py = ty * 16 + (p >> 4);
            T[p] = 1.f; done[p] = !(px < W && py < H); ndone += done[p];
          }
          // per-shape block state: done count per block evaluated lazily
          int j = 0;
          int stop_at[NSHAPE][32];   // record index at which each block became all-done (or n)
          for (int s = 0; s < NSHAPE; s++) for (int k = 0; k < 32; k++) stop_at[s][k] = -1;
          for (j = 0; j < n && ndone < 256; j++) {
            const int i = L[j].id;
            const float x = xy[2 * i], y = xy[2 * i + 1], A = co[4 * i], B = co[4 * i + 1], C = co[4 * i + 2], o = co[4 * i + 3];
            const float t = thr[i];
            const int tile_ok = in_ref_rect(x, y, (float)radii[i], tx, ty, gx, gy) &&
                                block_may(x, y, A, B, C, t, tx * 16.f, ty * 16.f, fminf(tx * 16.f + 15, W - 1), fminf(ty * 16.f + 15, H - 1));
            if (!tile_ok) continue;
            loc[2] += 1;
            unsigned char valid[256];
            for (int p = 0; p < 256; p++) {
              valid[p] = 0;
              if (done[p]) continue;
              const float dx = x - (tx * 16 + (p & 15)), dy = y - (ty * 16 + (p >> 4));
              const float power = -0.5f * (A * dx * dx + C * dy * dy) - B * dx * dy;
              if (power > 0) continue;
              const float alpha = fminf(0.99f, o * expf(power));
              if (alpha < 1.f / 255.f) continue;
              const float tt = T[p] * (1 - alpha);
              if (tt < 1e-4f && !mode_bwd) { done[p] = 2; continue; }   // 2: becomes done after this record
              if (tt < 1e-4f && mode_bwd) { done[p] = 2; continue; }
              T[p] = tt; valid[p] = 1; loc[3] += 1;
            }
            for (int s = 0; s < NSHAPE; s++) {
              const int bw = SH_W[s], bh = SH_H[s], nbx = 16 / bw, nby = 16 / bh;
              for (int k = 0; k < nbx * nby; k++) {
                if (stop_at[s][k] >= 0) continue;
                const int bx0 = tx * 16 + (k % nbx) * bw, by0 = ty * 16 + (k / nbx) * bh;
                if (!block_may(x, y, A, B, C, t, (float)bx0, (float)by0, fminf(bx0 + bw - 1, W - 1), fminf(by0 + bh - 1, H - 1))) continue;
                loc[8 + 3 * s + 1] += 1;
                int useful = 0;
                for (int yy = 0; yy < bh; yy++) for (int xx = 0; xx < bw; xx++) useful += valid[((k / nbx) * bh + yy) * 16 + (k % nbx) * bw + xx];
                loc[8 + 3 * s + 2] += useful;
              }
            }
            for (int p = 0; p < 256; p++) if (done[p] == 2) { done[p] = 1; ndone++; }
            for (int s = 0; s < NSHAPE; s++) {
              const int bw = SH_W[s], bh = SH_H[s], nbx = 16 / bw, nby = 16 / bh;
              for (int k = 0; k < nbx * nby; k++) {
                if (stop_at[s][k] >= 0) continue;
                int all = 1;
                for (int yy = 0; yy < bh && all; yy++) for (int xx = 0; xx < bw; xx++) if (!done[((k / nbx) * bh + yy) * 16 + (k % nbx) * bw + xx]) { all = 0; break; }
                if (all) stop_at[s][k] = j + 1;
              }
            }
          }
          loc[4] += j;
          loc[1] += (j + 255) / 256;
          for (int s = 0; s < NSHAPE; s++) {
            const int nblk = (16 / SH_W[s]) * (16 / SH_H[s]);
            for (int k = 0; k < nblk; k++) {
              const int st = stop_at[s][k] >= 0 ? stop_at[s][k] : j;
              loc[8 + 3 * s + 0] += (st + 31) / 32;
            }
          }
        }
#pragma omp critical
        for (int k = 0; k < 64; k++) acc[k] += loc[k];
      }
      for (int k = 1; k < 64; k++) out[k] = acc[k];
      free(ents); free(start); free(fill);
    }
  }
  free(cnt); free(thr);
  return 0;
}
