// gta_host.cu -- the reference's C entry points (include/delaunay_tri.h, include/gta_tri.h,
// include/gta_io.h) on top of the batch C-ABI of gta_b200.cu.
//
// Mirrors, call for call, what the reference does around its hot path:
//   dtinit / dtriangulate                    src/delaunay_tri.c:409-507
//   tessellate_area / delaunay_tessellate    src/gta_tri.c:41-78, :80-329
//   delaunay_surface_area                    src/gta_tri.c:332-413
//   print_areas / free_tri_area              src/gta_tri.c:448-473, :500-504
//   init_log / print_log / close_log         extern/gkut/src/gkut_log.c:21-50
//   read_traj / ndx_filter_traj              extern/gkut/src/gkut_io.c:20-97 (formats decoded here, no GROMACS)
// None of these functions computes a triangulation or an area on the CPU: they stage buffers and call
// gta_b200_tessellate_*; without a CUDA device they report the library error and produce no result.
#include <cuda_runtime.h>
#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>
#include "../../include/gta_b200.h"
#include "../../include/delaunay_tri.h"
#include "../../include/gta_tri.h"
#include "../../include/gta_io.h"

namespace {
thread_local int g_dt_status = 0;
thread_local int g_call_rc = 0;                 // GTA_B200_* return code behind the last kept (void) entry point on this thread
thread_local std::vector<int> g_frame_status;
const double kNaN = __builtin_nan("");
FILE* g_log = nullptr;

int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------------ logging
void init_log(const char* logfile, int argc, char* argv[]) {
    g_log = fopen(logfile, "a");
    if (!g_log) return;
    time_t t = time(nullptr);
    struct tm* lt = localtime(&t);
    fprintf(g_log, "\n");
    for (int i = 0; i < argc; ++i) fprintf(g_log, "%s ", argv[i]);
    fprintf(g_log, "\nRun: %d-%d-%d %d:%d:%d\n", lt->tm_mon + 1, lt->tm_mday, lt->tm_year + 1900, lt->tm_hour, lt->tm_min, lt->tm_sec);
}
void close_log(void) { if (g_log) fclose(g_log); g_log = nullptr; }
void print_log(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt); vprintf(fmt, ap); va_end(ap);
    if (g_log) { va_start(ap, fmt); vfprintf(g_log, fmt, ap); va_end(ap); }
}

// ------------------------------------------------------------------------------------------ triangulator
void dtinit(void) {
    if (gta_b200_device_count() < 1)
        fprintf(stderr, "TRIANGULATION ERROR: no CUDA device; libgtessla_b200 has no CPU implementation.\n");
}
int dt_last_status(void) { return g_dt_status; }

void dtriangulate(struct dTriangulation* tri) {
    g_dt_status = 0; g_call_rc = GTA_B200_OK;
    if (tri->npoints < 2) {                                   // reference delaunay_tri.c:416-421
        fprintf(stderr, "TRIANGULATION ERROR: Only %d points? That's not enough!\n", tri->npoints);   // the reference's text
        g_dt_status = GTA_B200_ST_TOO_FEW_POINTS;
        return;
    }
    const int n = tri->npoints, cap = 2 * n;                  // at most 2(n-1)-2 triangles (delaunay_tri.c:608)
    int* t = (int*)malloc(sizeof(int) * 3 * (size_t)cap);
    int ntri = 0, status = 0;
    int rc;
    if (n <= gta_b200_max_points())
        rc = gta_b200_tessellate_batch(tri->points, 2, nullptr, 0, 1, n, 0.0, 0u, -1, 1,
                                       nullptr, nullptr, nullptr, &status, t, cap, &ntri, nullptr, 0, nullptr);
    else
        rc = gta_b200_triangulate_large(tri->points, n, -1, t, cap, &ntri, &status);
    if (rc != GTA_B200_OK) {
        fprintf(stderr, "TRIANGULATION ERROR: %s\n", gta_b200_last_error());
        g_dt_status = GTA_B200_ST_INTERNAL; g_call_rc = rc; free(t); return;
    }
    g_dt_status = status;
    if (status != 0) {
        if (status & GTA_B200_ST_DUPLICATE)
            fprintf(stderr, "TRIANGULATION ERROR: duplicate points (closer than 1e-12 in x and y) are not supported.\n");
        else
            fprintf(stderr, "TRIANGULATION ERROR: input rejected (status %d, see gta_b200.h).\n", status);
        free(t); return;                                       // struct left untouched, as on the reference's error paths
    }
    tri->triangles = (int*)realloc(t, sizeof(int) * 3 * (size_t)(ntri > 0 ? ntri : 1));
    tri->ntriangles = ntri;
    tri->nverts = n;
}

// ------------------------------------------------------------------------------------------ area pipeline
const int* gta_last_frame_status(void) { return g_frame_status.empty() ? nullptr : g_frame_status.data(); }
int gta_last_call_rc(void) { return g_call_rc; }

static void write_showme(const char* node, const char* ele, const double* xy, int npts, const int* tri, int ntri) {
    // reference print_dtrifiles, src/gta_tri.c:475-498
    FILE* f = fopen(node, "w");
    if (f) {
        fprintf(f, "%d\t2\t0\t0\n", npts);
        for (int i = 0; i < npts; ++i) fprintf(f, "%d\t%f\t%f\n", i, xy[2 * i], xy[2 * i + 1]);
        fclose(f);
    }
    f = fopen(ele, "w");
    if (f) {
        fprintf(f, "%d\t3\t0\n", ntri);
        for (int i = 0; i < ntri; ++i) fprintf(f, "%d\t%d\t%d\t%d\n", i, tri[3 * i], tri[3 * i + 1], tri[3 * i + 2]);
        fclose(f);
    }
}

void delaunay_tessellate(rvec** x, matrix* box, real espace, int nthreads, struct tri_area* areas, unsigned char flags) {
    const int F = areas->nframes, N = areas->natoms;
    g_call_rc = GTA_B200_OK;
    if (nthreads > 1 || nthreads <= 0) print_log("Triangulation will be parallelized.\n");   // gta_tri.c:93-94
    dtinit();
    areas->area = (real*)calloc((size_t)(F > 0 ? F : 1), sizeof(real));
    areas->area2Dbox = (real*)calloc((size_t)(F > 0 ? F : 1), sizeof(real));
    if (flags & GTA_2D) areas->area2D = (real*)calloc((size_t)(F > 0 ? F : 1), sizeof(real));
    if (flags & GTA_CORRECT) print_log("Triangulating and correcting %d frames...\n", F);  // gta_tri.c:104
    else print_log("Triangulating %d frames...\n", F);                                        // gta_tri.c:299
    g_frame_status.assign((size_t)(F > 0 ? F : 1), 0);
    if (F <= 0) return;
    const unsigned fl = flags & (GTA_CORRECT | GTA_2D);
    const int maxp = gta_b200_max_points_for(&box[0][0][0], 9, F, N, espace, fl);
    int rc;
    std::string large_msg;                                     // (an error of a worker thread of the large-frame branch)
    if (maxp > gta_b200_max_points()) {
        // frames too large for the batched shared-memory kernels (e.g. a whole system without -n): one frame at a time
        // through the large-frame path (its own correction-point stage, the same corr_points as the batched kernels)
        // Frames are independent (gta_tri.c:106): with several GPUs every device takes every ng-th frame, one host thread each.
        const int want = env_int("GTA_B200_NGPUS", 0);
        const int ng = std::max(1, std::min(F, want <= 0 ? gta_b200_device_count() : std::min(want, gta_b200_device_count())));
        std::vector<int> rcs((size_t)ng, GTA_B200_OK);
        std::vector<std::string> msgs((size_t)ng);
        int* const fstat = g_frame_status.data();                  // (thread_local: the workers must write the caller's)
        auto work = [&, fstat](int g) {
            for (int fr = g; fr < F; fr += ng) {
                double a3 = 0.0, a2 = 0.0, ab = 0.0; int st = 0;
                const int r = gta_b200_tessellate_large(&x[fr][0][0], N, &box[fr][0][0], espace, fl, ng > 1 ? g : -1, &a3, &a2, &ab, nullptr, nullptr, &st);
                if (r != GTA_B200_OK) { rcs[(size_t)g] = r; msgs[(size_t)g] = gta_b200_last_error(); return; }     // (the message is this thread's)
                fstat[fr] = st;
                areas->area[fr] = st ? kNaN : a3;
                if (areas->area2D) areas->area2D[fr] = st ? kNaN : a2;
                areas->area2Dbox[fr] = ab;
            }
        };
        if (ng == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (int g = 0; g < ng; ++g) th.emplace_back(work, g);
            for (auto& t : th) t.join();
        }
        rc = GTA_B200_OK;
        for (int g = 0; g < ng; ++g) if (rcs[(size_t)g] != GTA_B200_OK && rc == GTA_B200_OK) { rc = rcs[(size_t)g]; large_msg = msgs[(size_t)g]; }
    } else if (flags & GTA_PRINT) {
        // one frame at a time (the reference also serialises -print, g_tessla.c:116); files are numbered from 1
        static int iter = 0;
        std::vector<int> tri(3 * (size_t)(2 * maxp)); std::vector<double> corr(3 * (size_t)(maxp - N + 1)), xy(2 * (size_t)maxp);
        rc = GTA_B200_OK;
        for (int fr = 0; fr < F && rc == GTA_B200_OK; ++fr) {
            int ntri = 0, ncorr = 0;
            rc = gta_b200_tessellate_batch(&x[fr][0][0], 3, &box[fr][0][0], 9, 1, N, espace, fl, -1, 1,
                                           areas->area + fr, areas->area2D ? areas->area2D + fr : nullptr, areas->area2Dbox + fr,
                                           &g_frame_status[fr], tri.data(), 2 * maxp, &ntri, corr.data(), maxp - N + 1, &ncorr);
            if (rc != GTA_B200_OK) break;
            for (int i = 0; i < N; ++i) { xy[2 * i] = x[fr][i][XX]; xy[2 * i + 1] = x[fr][i][YY]; }
            for (int i = 0; i < ncorr; ++i) { xy[2 * (N + i)] = corr[3 * i]; xy[2 * (N + i) + 1] = corr[3 * i + 1]; }
            char n1[64], n2[64];
            ++iter;
            snprintf(n1, sizeof n1, "triangles%d.node", iter); snprintf(n2, sizeof n2, "triangles%d.ele", iter);
            write_showme(n1, n2, xy.data(), N + ncorr, tri.data(), ntri);
        }
    } else {
        const int ngpus = env_int("GTA_B200_NGPUS", 0);
        rc = gta_b200_tessellate_frames((const double* const*)x, &box[0][0][0], F, N, espace, fl, ngpus <= 0 ? gta_b200_device_count() : ngpus,
                                        nthreads, areas->area, areas->area2D, areas->area2Dbox, g_frame_status.data());
    }
    if (rc != GTA_B200_OK) {
        // every function of this interface is void (reference include/gta_tri.h): the failure is reported on the log, in
        // gta_last_call_rc(), and as NaN areas with an INTERNAL status on every frame -- never as a table of zeros
        print_log("TESSELLATION ERROR: %s\n", large_msg.empty() ? gta_b200_last_error() : large_msg.c_str());
        g_call_rc = rc;
        for (int fr = 0; fr < F; ++fr) {
            areas->area[fr] = kNaN; if (areas->area2D) areas->area2D[fr] = kNaN;
            g_frame_status[fr] |= GTA_B200_ST_INTERNAL;
        }
        return;
    }
    int bad = 0;
    for (int fr = 0; fr < F; ++fr) bad += g_frame_status[fr] != 0;
    if (bad) print_log("WARNING: %d frame(s) were rejected (NaN area); see gta_last_frame_status().\n", bad);
}

void delaunay_surface_area(const rvec* x, matrix box, int natoms, unsigned char flags, real* a2D, real* a3D) {
    (void)box;
    static int iter = 0;                                       // numbers the -print files, as in the reference (gta_tri.c:338)
    g_dt_status = 0; g_call_rc = GTA_B200_OK;
    ++iter;
    const bool print = (flags & GTA_PRINT) != 0;
    if (!a2D && !a3D && !print) return;
    double a3 = 0.0, a2 = 0.0; int st = 0, ntri = 0, rc;
    std::vector<int> tri;
    if (natoms > gta_b200_max_points()) {
        rc = gta_b200_surface_area_large(&x[0][0], natoms, a2D ? GTA_B200_2D : 0u, -1, &a3, &a2, &ntri, &st);
        if (rc == GTA_B200_OK && print && st == 0) {
            std::vector<double> xy2(2 * (size_t)natoms);
            for (int i = 0; i < natoms; ++i) { xy2[2 * i] = x[i][XX]; xy2[2 * i + 1] = x[i][YY]; }
            tri.resize(3 * (size_t)(ntri > 0 ? ntri : 1));
            rc = gta_b200_triangulate_large(xy2.data(), natoms, -1, tri.data(), ntri, &ntri, &st);
        }
    } else {
        if (print) tri.resize(3 * (size_t)(2 * (natoms > 1 ? natoms : 1)));
        rc = gta_b200_tessellate_batch(&x[0][0], 3, nullptr, 0, 1, natoms, 0.0, a2D ? GTA_B200_2D : 0u, -1, 1,
                                       &a3, a2D ? &a2 : nullptr, nullptr, &st, print ? tri.data() : nullptr, print ? 2 * natoms : 0, &ntri,
                                       nullptr, 0, nullptr);
    }
    if (rc != GTA_B200_OK) {
        fprintf(stderr, "TRIANGULATION ERROR: %s\n", gta_b200_last_error());
        g_dt_status = GTA_B200_ST_INTERNAL; g_call_rc = rc;
        if (a3D) *a3D = kNaN;
        if (a2D) *a2D = kNaN;
        return;
    }
    g_dt_status = st;
    if (st & GTA_B200_ST_TOO_FEW_POINTS) fprintf(stderr, "TRIANGULATION ERROR: Only %d points? That's not enough!\n", natoms);
    if (print && st == 0) {                                    // triangles<iter>.node / .ele (gta_tri.c:355-360)
        std::vector<double> xy(2 * (size_t)natoms);
        for (int i = 0; i < natoms; ++i) { xy[2 * i] = x[i][XX]; xy[2 * i + 1] = x[i][YY]; }
        char n1[64], n2[64];
        snprintf(n1, sizeof n1, "triangles%d.node", iter); snprintf(n2, sizeof n2, "triangles%d.ele", iter);
        write_showme(n1, n2, xy.data(), natoms, tri.data(), ntri);
    }
    if (a3D) *a3D = (st & ~GTA_B200_ST_TOO_FEW_POINTS) ? kNaN : a3;
    if (a2D) *a2D = (st & ~GTA_B200_ST_TOO_FEW_POINTS) ? kNaN : a2;
}

void print_areas(const char* fname, const struct tri_area* areas) {
    FILE* f = fopen(fname, "w");
    if (!f) { print_log("Could not open %s for writing.\n", fname); return; }
    real sum = 0;
    if (areas->area2D) {
        fprintf(f, "# FRAME\tAREA\t2DAREA\tBOX-AREA\t\"\"/PARTICLE\n");
        for (int i = 0; i < areas->nframes; ++i) {
            fprintf(f, "%d\t%f\t%f\t%f\t%f\t%f\t%f\n", i, areas->area[i], areas->area2D[i], areas->area2Dbox[i],
                    areas->area[i] / areas->natoms, areas->area2D[i] / areas->natoms, areas->area2Dbox[i] / areas->natoms);
            sum += areas->area[i];
        }
    } else {
        fprintf(f, "# FRAME\tAREA\tBOX-AREA\t\"\"/PARTICLE\n");
        for (int i = 0; i < areas->nframes; ++i) {
            fprintf(f, "%d\t%f\t%f\t%f\t%f\n", i, areas->area[i], areas->area2Dbox[i],
                    areas->area[i] / areas->natoms, areas->area2Dbox[i] / areas->natoms);
            sum += areas->area[i];
        }
    }
    print_log("Average surface area: %f\n", sum / areas->nframes);
    print_log("Average area per particle: %f\n", (sum / areas->nframes) / areas->natoms);
    fclose(f);
    print_log("Surface areas saved to %s\n", fname);
}

void free_tri_area(struct tri_area* areas) {
    free(areas->area); free(areas->area2D); free(areas->area2Dbox);
    areas->area = areas->area2D = areas->area2Dbox = nullptr;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------ trajectory input
namespace {

// A trajectory as the readers deliver it: ONE contiguous [nframes][nkeep][3] array (what the batch entry points want;
// the reference's layout -- one heap block per frame, extern/gkut/src/gkut_io.c:37 -- is made from it only for the kept
// read_traj). `keep` (optional) is the index group: the filter is applied while reading, so the atoms that are not
// selected are never stored.
struct Traj {
    std::vector<double> xyz, box;                    // box: 9 per frame
    const std::vector<int>* keep = nullptr;
    int natoms = -1, nkeep = -1, nframes = 0;
    bool add(int n, const std::vector<double>& fr, const double b[9]) {
        if (natoms < 0) {
            natoms = n; nkeep = keep ? (int)keep->size() : n;
            if (keep) for (int a : *keep) if (a < 0 || a >= n) { fprintf(stderr, "index group: atom %d out of range (1..%d)\n", a + 1, n); return false; }
        }
        if (n != natoms) { fprintf(stderr, "read_traj: frame %d has %d atoms, expected %d\n", nframes, n, natoms); return false; }
        if (keep) for (int a : *keep) xyz.insert(xyz.end(), fr.begin() + 3 * (size_t)a, fr.begin() + 3 * (size_t)a + 3);
        else xyz.insert(xyz.end(), fr.begin(), fr.begin() + 3 * (size_t)n);
        box.insert(box.end(), b, b + 9);
        ++nframes;
        return true;
    }
};

bool ends_with(const std::string& s, const char* suf) {
    size_t n = strlen(suf);
    if (s.size() < n) return false;
    for (size_t i = 0; i < n; ++i) if (tolower((unsigned char)s[s.size() - n + i]) != suf[i]) return false;
    return true;
}

bool read_gro(FILE* f, Traj& t) {
    char line[512];
    std::vector<double> xyz;
    while (fgets(line, sizeof line, f)) {                     // title
        if (!fgets(line, sizeof line, f)) break;
        int n = atoi(line);
        if (n <= 0) return false;
        xyz.resize(3 * (size_t)n);
        for (int i = 0; i < n; ++i) {
            if (!fgets(line, sizeof line, f)) return false;
            if (strlen(line) <= 20) return false;                  // a truncated atom line
            // fixed columns: 20 characters of residue/atom fields, then x y z with a common width that is
            // inferred from the spacing of the decimal points (the .gro precision is not fixed at 8.3)
            const char* p = line + 20;
            const char* d1 = strchr(p, '.');
            const char* d2 = d1 ? strchr(d1 + 1, '.') : nullptr;
            int w = (d1 && d2) ? (int)(d2 - d1) : 8;
            char buf[64];
            for (int c = 0; c < 3; ++c) {
                if ((int)strlen(p) < w * (c + 1) - (c == 2 ? w - 1 : 0) && c < 2) return false;
                int len = w < 63 ? w : 63;
                strncpy(buf, p + c * w, len); buf[len] = 0;
                xyz[3 * (size_t)i + c] = atof(buf);
            }
        }
        if (!fgets(line, sizeof line, f)) return false;
        double v[9] = {0}, b[9] = {0};
        int k = sscanf(line, "%lf %lf %lf %lf %lf %lf %lf %lf %lf", v, v + 1, v + 2, v + 3, v + 4, v + 5, v + 6, v + 7, v + 8);
        if (k < 3) return false;
        // .gro order: v1(x) v2(y) v3(z) v1(y) v1(z) v2(x) v2(z) v3(x) v3(y)
        b[0] = v[0]; b[4] = v[1]; b[8] = v[2];
        if (k >= 9) { b[1] = v[3]; b[2] = v[4]; b[3] = v[5]; b[5] = v[6]; b[6] = v[7]; b[7] = v[8]; }
        if (!t.add(n, xyz, b)) return false;
    }
    return t.nframes > 0;
}

bool read_pdb(FILE* f, Traj& t) {
    char line[512];
    std::vector<double> xyz; double b[9] = {0};
    auto flush = [&]() -> bool {
        if (xyz.empty()) return true;
        bool ok = t.add((int)(xyz.size() / 3), xyz, b);
        xyz.clear();
        return ok;
    };
    auto field = [](const char* s, int a, int n) { char buf[32]; int len = (int)strlen(s); if (len < a) return 0.0; strncpy(buf, s + a, n); buf[n] = 0; return atof(buf); };
    while (fgets(line, sizeof line, f)) {
        if (!strncmp(line, "CRYST1", 6)) {
            memset(b, 0, sizeof b);                            // orthogonal part only (the hot path reads [0][0], [1][1])
            b[0] = field(line, 6, 9) / 10.0; b[4] = field(line, 15, 9) / 10.0; b[8] = field(line, 24, 9) / 10.0;
        } else if (!strncmp(line, "ATOM", 4) || !strncmp(line, "HETATM", 6)) {
            xyz.push_back(field(line, 30, 8) / 10.0); xyz.push_back(field(line, 38, 8) / 10.0); xyz.push_back(field(line, 46, 8) / 10.0);
        } else if (!strncmp(line, "ENDMDL", 6) || (!strncmp(line, "END", 3) && !isalpha((unsigned char)line[3]))) {
            if (!flush()) return false;
        }
    }
    return flush() && t.nframes > 0;
}

// XDR (big-endian) primitives for .trr
bool xdr_u32(FILE* f, uint32_t* v) { unsigned char b[4]; if (fread(b, 1, 4, f) != 4) return false; *v = (uint32_t)b[0] << 24 | (uint32_t)b[1] << 16 | (uint32_t)b[2] << 8 | b[3]; return true; }
bool xdr_i32(FILE* f, int* v) { uint32_t u; if (!xdr_u32(f, &u)) return false; *v = (int)u; return true; }
bool xdr_real(FILE* f, bool dbl, double* v) {
    if (dbl) { uint32_t hi, lo; if (!xdr_u32(f, &hi) || !xdr_u32(f, &lo)) return false; uint64_t u = (uint64_t)hi << 32 | lo; memcpy(v, &u, 8); return true; }
    uint32_t u; if (!xdr_u32(f, &u)) return false; float fl; memcpy(&fl, &u, 4); *v = fl; return true;
}

bool read_trr(FILE* f, Traj& t) {
    std::vector<double> xyz;
    for (;;) {
        int magic;
        if (!xdr_i32(f, &magic)) break;                        // clean EOF
        if (magic != 1993) return false;
        int slen, slen2;
        if (!xdr_i32(f, &slen) || !xdr_i32(f, &slen2)) return false;   // xdr_string: length word twice (count, then opaque length)
        if (slen2 < 0 || slen2 > 256) return false;
        if (fseek(f, (slen2 + 3) & ~3, SEEK_CUR)) return false;
        int sz[10], natoms, step, nre;
        for (int i = 0; i < 10; ++i) if (!xdr_i32(f, &sz[i])) return false;   // ir e box vir pres top sym x v f
        if (!xdr_i32(f, &natoms) || !xdr_i32(f, &step) || !xdr_i32(f, &nre)) return false;
        const int box_size = sz[2], x_size = sz[7];
        for (int i = 0; i < 10; ++i) if (sz[i] < 0) return false;           // counts come from the file: never seek backwards
        if (natoms < 0 || natoms > (1 << 28)) return false;
        bool dbl;
        if (box_size) dbl = box_size == 9 * 8; else if (natoms > 0 && x_size) dbl = x_size == natoms * 3 * 8; else return false;
        double tl[2];
        if (!xdr_real(f, dbl, &tl[0]) || !xdr_real(f, dbl, &tl[1])) return false;
        double b[9] = {0};
        if (box_size) for (int i = 0; i < 9; ++i) if (!xdr_real(f, dbl, &b[i])) return false;
        if (fseek(f, sz[3] + sz[4], SEEK_CUR)) return false;                   // virial, pressure
        if (!x_size) { if (fseek(f, sz[8] + sz[9], SEEK_CUR)) return false; continue; }   // frame without coordinates
        if ((long long)x_size != (long long)natoms * 3 * (dbl ? 8 : 4)) return false;
        xyz.resize(3 * (size_t)natoms);
        for (size_t i = 0; i < xyz.size(); ++i) if (!xdr_real(f, dbl, &xyz[i])) return false;
        if (fseek(f, sz[8] + sz[9], SEEK_CUR)) return false;                   // velocities, forces
        if (!t.add(natoms, xyz, b)) return false;
    }
    return t.nframes > 0;
}


// ---- .xtc: XDR frame header + GROMACS' compressed coordinates (the "xdr3dfcoord" scheme of libxdrf / xdrfile).
// Restated from the published description of that format (no GROMACS or xdrfile source is available in this
// build environment, and no GROMACS-written file to check against): positions are integers at 1/precision nm, coded
// either with the bit width of the frame's bounding box or, in runs, as small offsets from the previous atom in a
// mixed-radix number whose radix comes from the `magicints` table. tests/ round-trips it against an independent
// Python encoder of the same description.
const int xtc_magicints[] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812,
                             1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007, 32768, 41285, 52015,
                             65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561, 832255, 1048576, 1321122,
                             1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216};
const int XTC_FIRSTIDX = 9, XTC_LASTIDX = (int)(sizeof(xtc_magicints) / sizeof(int));

struct XtcBits {                      // MSB-first bit reader over the opaque block
    const unsigned char* p; size_t n, pos = 0; unsigned lastbits = 0, lastbyte = 0; bool bad = false;
    unsigned byte() { if (pos >= n) { bad = true; return 0; } return p[pos++]; }
    unsigned bits(int nbits) {
        const unsigned mask = nbits >= 32 ? 0xffffffffu : ((1u << nbits) - 1u);
        unsigned num = 0;
        while (nbits >= 8) { lastbyte = (lastbyte << 8) | byte(); num |= (lastbyte >> lastbits) << (nbits - 8); nbits -= 8; }
        if (nbits > 0) {
            if ((int)lastbits < nbits) { lastbits += 8; lastbyte = (lastbyte << 8) | byte(); }
            lastbits -= nbits;
            num |= (lastbyte >> lastbits) & ((1u << nbits) - 1u);
        }
        return num & mask;
    }
    void ints(int nbits, const unsigned sizes[3], int nums[3]) {      // mixed-radix number -> 3 digits
        unsigned bytes[32] = {0}; int nbytes = 0;
        while (nbits > 8) { bytes[nbytes++] = bits(8); nbits -= 8; }
        if (nbits > 0) bytes[nbytes++] = bits(nbits);
        for (int i = 2; i > 0; --i) {
            unsigned num = 0;
            for (int j = nbytes - 1; j >= 0; --j) { num = (num << 8) | bytes[j]; const unsigned q = num / sizes[i]; bytes[j] = q; num -= q * sizes[i]; }
            nums[i] = (int)num;
        }
        nums[0] = (int)(bytes[0] | (bytes[1] << 8) | (bytes[2] << 16) | (bytes[3] << 24));
    }
};
int xtc_sizeofint(unsigned size) { unsigned num = 1; int nb = 0; while (size >= num && nb < 32) { ++nb; num <<= 1; } return nb; }
int xtc_sizeofints(const unsigned sizes[3]) {
    unsigned bytes[32]; int nbytes = 1; bytes[0] = 1;
    for (int i = 0; i < 3; ++i) {
        unsigned tmp = 0; int bc = 0;
        for (; bc < nbytes; ++bc) { tmp = bytes[bc] * sizes[i] + tmp; bytes[bc] = tmp & 0xff; tmp >>= 8; }
        while (tmp != 0) { bytes[bc++] = tmp & 0xff; tmp >>= 8; }
        nbytes = bc;
    }
    unsigned num = 1; int nb = 0; --nbytes;
    while (bytes[nbytes] >= num) { ++nb; num *= 2; }
    return nb + nbytes * 8;
}
bool xdr_f32(FILE* f, double* v) { return xdr_real(f, false, v); }

bool read_xtc(FILE* f, Traj& t) {
    std::vector<double> xyz; std::vector<unsigned char> buf; std::vector<int> ip;
    for (;;) {
        int magic, natoms, step;
        if (!xdr_i32(f, &magic)) break;                        // clean EOF
        if (magic != 1995 || !xdr_i32(f, &natoms) || !xdr_i32(f, &step) || natoms < 0 || natoms > (1 << 28)) return false;
        double tm, b[9];
        if (!xdr_f32(f, &tm)) return false;
        for (int i = 0; i < 9; ++i) if (!xdr_f32(f, &b[i])) return false;
        int lsize;
        if (!xdr_i32(f, &lsize) || lsize != natoms) return false;
        xyz.assign(3 * (size_t)natoms, 0.0);
        if (lsize <= 9) {                                      // tiny frames are stored as plain floats
            for (size_t i = 0; i < xyz.size(); ++i) if (!xdr_f32(f, &xyz[i])) return false;
            if (!t.add(natoms, xyz, b)) return false;
            continue;
        }
        double precision;
        int minint[3], maxint[3], smallidx, nbytes;
        if (!xdr_f32(f, &precision) || !(precision > 0)) return false;
        for (int i = 0; i < 3; ++i) if (!xdr_i32(f, &minint[i])) return false;
        for (int i = 0; i < 3; ++i) if (!xdr_i32(f, &maxint[i])) return false;
        if (!xdr_i32(f, &smallidx) || smallidx < XTC_FIRSTIDX || smallidx >= XTC_LASTIDX) return false;
        if (!xdr_i32(f, &nbytes) || nbytes < 0 || nbytes > (1 << 30)) return false;
        buf.resize(((size_t)nbytes + 3) & ~(size_t)3);
        if (fread(buf.data(), 1, buf.size(), f) != buf.size()) return false;
        unsigned sizeint[3], sizesmall[3]; int bitsizeint[3] = {0, 0, 0}, bitsize;
        for (int i = 0; i < 3; ++i) { if (maxint[i] < minint[i]) return false; sizeint[i] = (unsigned)(maxint[i] - minint[i]) + 1u; }
        if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffffu) { for (int i = 0; i < 3; ++i) bitsizeint[i] = xtc_sizeofint(sizeint[i]); bitsize = 0; }
        else bitsize = xtc_sizeofints(sizeint);
        int smaller = xtc_magicints[smallidx - 1 > XTC_FIRSTIDX ? smallidx - 1 : XTC_FIRSTIDX] / 2;
        int smallnum = xtc_magicints[smallidx] / 2;
        sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)xtc_magicints[smallidx];
        XtcBits br; br.p = buf.data(); br.n = (size_t)nbytes;
        const float inv = 1.0f / (float)precision;                // the format's own arithmetic is single precision
        int run = 0, i = 0, prev[3], cur[3];
        size_t o = 0;
        while (i < lsize) {
            if (bitsize == 0) { for (int c = 0; c < 3; ++c) cur[c] = (int)br.bits(bitsizeint[c]); }
            else br.ints(bitsize, sizeint, cur);
            ++i;
            for (int c = 0; c < 3; ++c) { cur[c] += minint[c]; prev[c] = cur[c]; }
            const unsigned flag = br.bits(1);
            int is_smaller = 0;
            if (flag == 1) { run = (int)br.bits(5); is_smaller = run % 3; run -= is_smaller; --is_smaller; }
            if (run > 0) {
                if (i + run / 3 > lsize) return false;
                for (int k = 0; k < run; k += 3) {
                    br.ints(smallidx, sizesmall, cur);
                    ++i;
                    for (int c = 0; c < 3; ++c) cur[c] += prev[c] - smallnum;
                    if (k == 0) {
                        // the first atom of a run is stored after the second one (better compression of water): swap back
                        for (int c = 0; c < 3; ++c) { const int tmp = cur[c]; cur[c] = prev[c]; prev[c] = tmp; }
                        for (int c = 0; c < 3; ++c) xyz[o++] = (double)((float)prev[c] * inv);
                    } else for (int c = 0; c < 3; ++c) prev[c] = cur[c];
                    for (int c = 0; c < 3; ++c) xyz[o++] = (double)((float)cur[c] * inv);
                }
            } else for (int c = 0; c < 3; ++c) xyz[o++] = (double)((float)cur[c] * inv);
            smallidx += is_smaller;
            if (smallidx < XTC_FIRSTIDX || smallidx >= XTC_LASTIDX) return false;
            if (is_smaller < 0) { smallnum = smaller; smaller = smallidx > XTC_FIRSTIDX ? xtc_magicints[smallidx - 1] / 2 : 0; }
            else if (is_smaller > 0) { smaller = smallnum; smallnum = xtc_magicints[smallidx] / 2; }
            sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)xtc_magicints[smallidx];
            if (br.bad) return false;
        }
        if (o != xyz.size()) return false;
        if (!t.add(natoms, xyz, b)) return false;
    }
    return t.nframes > 0;
}

bool read_any(const char* traj_fname, Traj& t) {
    std::string name(traj_fname ? traj_fname : "");
    FILE* f = fopen(traj_fname, "rb");
    if (!f) { fprintf(stderr, "read_traj: cannot open %s\n", traj_fname); return false; }
    bool ok;
    try {                                                      // (nothing may propagate through the C interface)
        if (ends_with(name, ".gro")) ok = read_gro(f, t);
        else if (ends_with(name, ".pdb")) ok = read_pdb(f, t);
        else if (ends_with(name, ".trr")) ok = read_trr(f, t);
        else if (ends_with(name, ".xtc")) ok = read_xtc(f, t);
        else { fprintf(stderr, "read_traj: %s: unknown trajectory format (supported: .gro .pdb .trr .xtc)\n", traj_fname); fclose(f); return false; }
    } catch (...) { ok = false; }
    fclose(f);
    if (!ok) fprintf(stderr, "read_traj: %s: malformed or empty trajectory\n", traj_fname);
    return ok;
}

// first group of an index file, 0-based
bool read_ndx(const char* ndx_fname, std::vector<int>& idx) {
    FILE* f = fopen(ndx_fname, "r");
    if (!f) { fprintf(stderr, "ndx_filter_traj: cannot open %s\n", ndx_fname); return false; }
    char line[4096]; int groups = 0;
    while (fgets(line, sizeof line, f)) {
        const char* p = line; while (isspace((unsigned char)*p)) ++p;
        if (*p == '[') { if (++groups > 1) break; continue; }
        if (groups != 1) continue;
        char* q = (char*)p;
        for (;;) { char* end; long v = strtol(q, &end, 10); if (end == q) break; idx.push_back((int)v - 1); q = end; }
    }
    fclose(f);
    return true;
}

}  // namespace

extern "C" {

void read_traj(const char* traj_fname, rvec*** x, matrix** box, int* nframes, int* natoms, output_env_t* oenv) {
    (void)oenv;
    *x = nullptr; *box = nullptr; *nframes = 0; *natoms = 0;
    Traj t;
    if (!read_any(traj_fname, t)) return;
    // the reference's layout: one heap block per frame (extern/gkut/src/gkut_io.c:37)
    const int F = t.nframes, N = t.natoms;
    *x = (rvec**)malloc(sizeof(rvec*) * (size_t)F);
    *box = (matrix*)malloc(sizeof(matrix) * (size_t)F);
    for (int fr = 0; fr < F; ++fr) {
        (*x)[fr] = (rvec*)malloc(sizeof(rvec) * (size_t)(N > 0 ? N : 1));
        memcpy((*x)[fr], t.xyz.data() + 3 * (size_t)N * fr, sizeof(double) * 3 * (size_t)N);
    }
    memcpy(*box, t.box.data(), sizeof(double) * 9 * (size_t)F);
    *nframes = F; *natoms = N;
}

void ndx_filter_traj(const char* ndx_fname, rvec** pre_x, rvec*** new_x, int nframes, int* natoms) {
    std::vector<int> idx;
    read_ndx(ndx_fname, idx);
    const int n_in = *natoms, n = (int)idx.size();
    for (int i = 0; i < n; ++i)
        if (idx[i] < 0 || idx[i] >= n_in) {
            // a hard error, as with GROMACS: silently substituting a point would change the tessellation
            fprintf(stderr, "ndx_filter_traj: %s: atom %d out of range (1..%d)\n", ndx_fname, idx[i] + 1, n_in);
            *new_x = nullptr; *natoms = 0;
            return;
        }
    *new_x = (rvec**)malloc(sizeof(rvec*) * (size_t)(nframes > 0 ? nframes : 1));
    for (int fr = 0; fr < nframes; ++fr) {
        (*new_x)[fr] = (rvec*)malloc(sizeof(rvec) * (size_t)(n > 0 ? n : 1));
        for (int i = 0; i < n; ++i) {
            memcpy((*new_x)[fr][i], pre_x[fr][idx[i]], sizeof(rvec));
        }
    }
    *natoms = n;
}

void tessellate_area(const char* traj_fname, const char* ndx_fname, output_env_t* oenv, real espace, int nthreads,
                     struct tri_area* areas, unsigned char flags) {
    (void)oenv;
    areas->area = areas->area2D = areas->area2Dbox = nullptr;          // gta_tri.c:51-53
    areas->nframes = 0; areas->natoms = 0;
    g_call_rc = GTA_B200_OK;
    // The reference reads one heap block per frame, filters them into new blocks and hands the pointers on (gta_tri.c:55-70).
    // Here the reader streams the selected atoms of every frame into ONE contiguous array, which is what the batch entry
    // point wants: no per-frame allocation, no filter copy, no gather into the staging buffers.
    std::vector<int> idx;
    Traj t;
    if (ndx_fname) { if (!read_ndx(ndx_fname, idx)) { g_call_rc = GTA_B200_E_ARG; return; } t.keep = &idx; }
    if (!read_any(traj_fname, t)) { g_call_rc = GTA_B200_E_ARG; return; }
    const int F = t.nframes, N = t.nkeep;
    areas->nframes = F; areas->natoms = N;
    const unsigned fl = flags & (GTA_CORRECT | GTA_2D);
    const int maxp = gta_b200_max_points_for(t.box.data(), 9, F, N, espace, fl);
    if ((flags & GTA_PRINT) || maxp > gta_b200_max_points()) {
        // -print (one frame at a time) and frames beyond the batched kernel: through delaunay_tessellate on per-frame views
        std::vector<rvec*> x((size_t)F);
        for (int fr = 0; fr < F; ++fr) x[fr] = (rvec*)(t.xyz.data() + 3 * (size_t)N * fr);
        delaunay_tessellate(x.data(), (matrix*)t.box.data(), espace, nthreads, areas, flags);
        return;
    }
    if (nthreads > 1 || nthreads <= 0) print_log("Triangulation will be parallelized.\n");   // gta_tri.c:93-94
    dtinit();
    areas->area = (real*)calloc((size_t)(F > 0 ? F : 1), sizeof(real));
    areas->area2Dbox = (real*)calloc((size_t)(F > 0 ? F : 1), sizeof(real));
    if (flags & GTA_2D) areas->area2D = (real*)calloc((size_t)(F > 0 ? F : 1), sizeof(real));
    if (flags & GTA_CORRECT) print_log("Triangulating and correcting %d frames...\n", F);  // gta_tri.c:104
    else print_log("Triangulating %d frames...\n", F);                                        // gta_tri.c:299
    g_frame_status.assign((size_t)(F > 0 ? F : 1), 0);
    const int ngpus = env_int("GTA_B200_NGPUS", 0);
    const int rc = gta_b200_tessellate_multi(t.xyz.data(), 3, t.box.data(), 9, F, N, espace, fl, ngpus <= 0 ? gta_b200_device_count() : ngpus,
                                             nthreads > 0 ? nthreads : 0, areas->area, areas->area2D, areas->area2Dbox, g_frame_status.data());
    if (rc != GTA_B200_OK) {
        print_log("TESSELLATION ERROR: %s\n", gta_b200_last_error());
        g_call_rc = rc;
        for (int fr = 0; fr < F; ++fr) { areas->area[fr] = kNaN; if (areas->area2D) areas->area2D[fr] = kNaN; g_frame_status[fr] |= GTA_B200_ST_INTERNAL; }
        return;
    }
    int bad = 0;
    for (int fr = 0; fr < F; ++fr) bad += g_frame_status[fr] != 0;
    if (bad) print_log("WARNING: %d frame(s) were rejected (NaN area); see gta_last_frame_status().\n", bad);
}

}  // extern "C"
