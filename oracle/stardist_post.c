/* stardist_post.c -- CPU ORACLE (test infrastructure only) for the post-processing of
 * StarDist2D.predict_instances (improved_detection.py:63): dist_to_coord, the greedy polygon
 * non-maximum suppression and polygons_to_label.
 *
 * PARITY UNPINNED: stardist (and scikit-image, whose skimage.draw.polygon renders the labels) are
 * third-party packages absent from /root/reference and from this image; the reference pins no version
 * (README: unversioned pip install) and ships no golden vectors.  This file restates
 *   - stardist.geometry.geom2d.dist_to_coord / polygons_to_label / polygons_to_label_coord,
 *   - stardist.nms.non_maximum_suppression_sparse + lib/stardist2d_impl.cpp (greedy suppression in
 *     descending probability; a pair overlaps when intersection area / smaller area > nms_thresh),
 *   - skimage/_shared/_geometry + skimage/draw/_draw.pyx::_polygon and _pnpoly.h::point_in_polygon,
 * from their published sources as remembered.  One stated deviation: stardist intersects the polygons
 * with the Clipper library on integer-snapped vertices; here the intersection area is exact on the
 * float32 vertices (star-convex polygons as triangle fans, pairwise Sutherland-Hodgman clipping in
 * fp64).  Compile with -ffp-contract=off: the CUDA path uses the same statements with explicit
 * round-to-nearest operations, so decisions and labels agree bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NR 32

/* coord = (dist[:, None] * [sin, cos]).astype(float32); coord += points  (float32 += int64: the sum is
 * taken in float64 and stored as float32); area by the shoelace sum */
void sd_polygons(int n, const float* dist /* [n][32] */, const int32_t* pyx /* [n][2] */, const double* rsin,
                 const double* rcos, float* vy, float* vx, double* area) {
    for (int r = 0; r < n; ++r) {
        for (int k = 0; k < NR; ++k) {
            const float ty = (float)((double)dist[r * NR + k] * rsin[k]);
            const float tx = (float)((double)dist[r * NR + k] * rcos[k]);
            vy[r * NR + k] = (float)((double)ty + (double)pyx[2 * r]);
            vx[r * NR + k] = (float)((double)tx + (double)pyx[2 * r + 1]);
        }
        double s = 0.0;
        for (int k = 0; k < NR; ++k) {
            const int k1 = (k + 1) % NR;
            s = s + ((double)vx[r * NR + k] * (double)vy[r * NR + k1] - (double)vx[r * NR + k1] * (double)vy[r * NR + k]);
        }
        area[r] = 0.5 * fabs(s);
    }
}

static double cross_(double ax, double ay, double bx, double by, double px, double py) {
    return (bx - ax) * (py - ay) - (by - ay) * (px - ax);
}

static double tri_tri_area(const double* sx, const double* sy, const double* cx, const double* cy) {
    double mn1 = fmin(fmin(sx[0], sx[1]), sx[2]), mx1 = fmax(fmax(sx[0], sx[1]), sx[2]);
    double mn2 = fmin(fmin(cx[0], cx[1]), cx[2]), mx2 = fmax(fmax(cx[0], cx[1]), cx[2]);
    if (mx1 < mn2 || mx2 < mn1) return 0.0;
    mn1 = fmin(fmin(sy[0], sy[1]), sy[2]); mx1 = fmax(fmax(sy[0], sy[1]), sy[2]);
    mn2 = fmin(fmin(cy[0], cy[1]), cy[2]); mx2 = fmax(fmax(cy[0], cy[1]), cy[2]);
    if (mx1 < mn2 || mx2 < mn1) return 0.0;
    /* separating edge (both triangles are counter-clockwise): exactly 0 */
    for (int e = 0; e < 3; ++e) {
        const int e1 = e == 2 ? 0 : e + 1;
        if (cross_(cx[e], cy[e], cx[e1], cy[e1], sx[0], sy[0]) < 0.0 && cross_(cx[e], cy[e], cx[e1], cy[e1], sx[1], sy[1]) < 0.0 &&
            cross_(cx[e], cy[e], cx[e1], cy[e1], sx[2], sy[2]) < 0.0) return 0.0;
        if (cross_(sx[e], sy[e], sx[e1], sy[e1], cx[0], cy[0]) < 0.0 && cross_(sx[e], sy[e], sx[e1], sy[e1], cx[1], cy[1]) < 0.0 &&
            cross_(sx[e], sy[e], sx[e1], sy[e1], cx[2], cy[2]) < 0.0) return 0.0;
    }
    double px[8], py[8], qx[8], qy[8];
    int n = 3;
    for (int k = 0; k < 3; ++k) { px[k] = sx[k]; py[k] = sy[k]; }
    for (int e = 0; e < 3; ++e) {
        const double ax = cx[e], ay = cy[e], bx = cx[(e + 1) % 3], by = cy[(e + 1) % 3];
        int m = 0;
        for (int k = 0; k < n; ++k) {
            const int k1 = k + 1 == n ? 0 : k + 1;
            const double dc = cross_(ax, ay, bx, by, px[k], py[k]);
            const double dn = cross_(ax, ay, bx, by, px[k1], py[k1]);
            if (dc >= 0.0) { qx[m] = px[k]; qy[m] = py[k]; ++m; }
            if ((dc >= 0.0) != (dn >= 0.0)) {
                const double t = dc / (dc - dn);
                qx[m] = px[k] + t * (px[k1] - px[k]);
                qy[m] = py[k] + t * (py[k1] - py[k]);
                ++m;
            }
        }
        n = m;
        if (n == 0) return 0.0;
        for (int k = 0; k < n; ++k) { px[k] = qx[k]; py[k] = qy[k]; }
    }
    double s = 0.0;
    for (int k = 0; k < n; ++k) {
        const int k1 = k + 1 == n ? 0 : k + 1;
        s = s + (px[k] * py[k1] - px[k1] * py[k]);
    }
    return 0.5 * fabs(s);
}

/* intersection area / (smaller area + 1e-10) of polygons w and i */
double sd_overlap(const float* vy, const float* vx, const int32_t* pyx, const double* area, int w, int i) {
    float y0 = vy[w * NR], y1 = y0, x0 = vx[w * NR], x1 = x0, u0 = vy[i * NR], u1 = u0, v0 = vx[i * NR], v1 = v0;
    for (int k = 1; k < NR; ++k) {
        y0 = fminf(y0, vy[w * NR + k]); y1 = fmaxf(y1, vy[w * NR + k]);
        x0 = fminf(x0, vx[w * NR + k]); x1 = fmaxf(x1, vx[w * NR + k]);
        u0 = fminf(u0, vy[i * NR + k]); u1 = fmaxf(u1, vy[i * NR + k]);
        v0 = fminf(v0, vx[i * NR + k]); v1 = fmaxf(v1, vx[i * NR + k]);
    }
    if (y1 < u0 || u1 < y0 || x1 < v0 || v1 < x0) return 0.0;
    double inter = 0.0;
    for (int a = 0; a < NR; ++a) {
        const int a1 = (a + 1) % NR;
        const double sx[3] = {(double)pyx[2 * w + 1], (double)vx[w * NR + a], (double)vx[w * NR + a1]};
        const double sy[3] = {(double)pyx[2 * w], (double)vy[w * NR + a], (double)vy[w * NR + a1]};
        double part = 0.0;
        for (int st = 0; st < NR; ++st) {           /* the CUDA lane's order: b = a, a + 1, ... */
            const int b = (a + st) % NR;
            const int b1 = (b + 1) % NR;
            const double cx[3] = {(double)pyx[2 * i + 1], (double)vx[i * NR + b], (double)vx[i * NR + b1]};
            const double cy[3] = {(double)pyx[2 * i], (double)vy[i * NR + b], (double)vy[i * NR + b1]};
            part = part + tri_tri_area(sx, sy, cx, cy);
        }
        inter = inter + part;
    }
    return inter / (fmin(area[w], area[i]) + 1e-10);
}

/* greedy suppression over polygons already sorted by descending probability; keep[r] = 1 for survivors */
void sd_nms(int n, const float* vy, const float* vx, const int32_t* pyx, const double* area, double thr,
            uint8_t* keep) {
    uint8_t* sup = (uint8_t*)calloc((size_t)n + 1, 1);
    for (int i = 0; i < n; ++i) {
        keep[i] = !sup[i];
        if (sup[i]) continue;
        for (int j = i + 1; j < n; ++j) {
            if (sup[j]) continue;
            if (sd_overlap(vy, vx, pyx, area, i, j) > thr) sup[j] = 1;
        }
    }
    free(sup);
}

/* skimage _pnpoly.h point_in_polygon: 0 outside, 1 inside, 2 on an edge, 3 on a vertex */
static int pnpoly(int nv, const double* xp, const double* yp, double x, double y) {
    int l_cross = 0, r_cross = 0;
    const double eps = 1e-12;
    double x1 = xp[nv - 1] - x, y1 = yp[nv - 1] - y;
    for (int i = 0; i < nv; ++i) {
        const double x0 = xp[i] - x, y0 = yp[i] - y;
        if (-eps < x0 && x0 < eps && -eps < y0 && y0 < eps) return 3;
        if ((y0 > 0) != (y1 > 0)) {
            if (((x0 * y1 - x1 * y0) / (y1 - y0)) > 0) ++r_cross;
        }
        if ((y0 < 0) != (y1 < 0)) {
            if (((x0 * y1 - x1 * y0) / (y1 - y0)) < 0) ++l_cross;
        }
        x1 = x0; y1 = y0;
    }
    if ((r_cross & 1) != (l_cross & 1)) return 2;
    return r_cross & 1;
}
int sd_pnpoly(int nv, const double* xp, const double* yp, double x, double y) { return pnpoly(nv, xp, yp, x, y); }

/* polygons_to_label: the kept polygons (descending probability, label = index + 1) are drawn in ASCENDING
 * probability with skimage.draw.polygon(r, c, shape), later polygons overwriting earlier ones */
void sd_render(int n_kept, const float* vy, const float* vx, int H, int W, int32_t* labels) {
    memset(labels, 0, (size_t)H * W * sizeof(int32_t));
    for (int k = n_kept - 1; k >= 0; --k) {
        double yp[NR], xp[NR];
        float y0 = vy[k * NR], y1 = y0, x0 = vx[k * NR], x1 = x0;
        for (int j = 0; j < NR; ++j) {
            yp[j] = (double)vy[k * NR + j]; xp[j] = (double)vx[k * NR + j];
            y0 = fminf(y0, vy[k * NR + j]); y1 = fmaxf(y1, vy[k * NR + j]);
            x0 = fminf(x0, vx[k * NR + j]); x1 = fmaxf(x1, vx[k * NR + j]);
        }
        int minr = (int)fmaxf(0.f, y0), maxr = (int)ceilf(y1), minc = (int)fmaxf(0.f, x0), maxc = (int)ceilf(x1);
        if (maxr > H - 1) maxr = H - 1;
        if (maxc > W - 1) maxc = W - 1;
        for (int r = minr; r <= maxr; ++r)
            for (int c = minc; c <= maxc; ++c)
                if (pnpoly(NR, xp, yp, (double)c, (double)r)) labels[(size_t)r * W + c] = k + 1;
    }
}
