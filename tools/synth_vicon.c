/*
 * Synthetic Vicon Nexus CSV generator (bench/test infrastructure, not product code).
 *
 * Produces files of the layout the reference loader understands (SURVEY.md Appendix A/F,
 * modelled on /root/reference/sample_data/abridged_data.csv): a "Devices" section
 * (P force plates x 9 columns + one EMG device with E muscles, f_emg Hz) followed by an
 * all-commas separator row and a "Trajectories" section (M markers x 3 columns, f_traj Hz).
 * Every line is padded with trailing commas to the widest row, like a real export.
 *
 * Number formatting mimics the sample file: "%.6g", "%.2E" for 0<|v|<1e-4, bare integers
 * for centre-of-pressure columns and exact "0" for an unloaded plate; occluded markers are
 * three empty fields.  The vertical ground reaction follows a 4-trecho walking pattern with
 * >= 40 one-leg/two-leg alternations so that the segmenter has something to find, plus
 * short "chatter" flips (< min_phase_size rows) that it must ignore.
 *
 * Every row is generated from an RNG keyed by (seed, section, row), so the output does not
 * depend on the number of OpenMP threads.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC tools/synth_vicon.c -o tools/libms_synth.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint64_t s;
} rng_t;

static inline uint64_t splitmix64(uint64_t *x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double rng_uniform(rng_t *r) { return (double)(splitmix64(&r->s) >> 11) * (1.0 / 9007199254740992.0); }
static inline double rng_normal(rng_t *r) {
    double u1 = rng_uniform(r), u2 = rng_uniform(r);
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}
static inline rng_t rng_for(uint64_t seed, uint64_t section, uint64_t row) {
    rng_t r;
    uint64_t x = seed * 0xD1342543DE82EF95ull + section * 0xA24BAED4963EE407ull + row * 0x9FB21C651E98DF25ull + 12345;
    splitmix64(&x);
    r.s = x;
    return r;
}

typedef struct {
    int n_plates, n_emg, n_markers, f_emg, f_traj, crlf, width, trailing_blank;
    double seconds, blank_marker_frac;
    int64_t n_dev_rows, n_traj_rows;
    int subframes;
    uint64_t seed;
} cfg_t;

/* formatting ------------------------------------------------------------------------- */
static inline char *put_str(char *p, const char *s) {
    size_t n = strlen(s);
    memcpy(p, s, n);
    return p + n;
}
static inline char *put_int(char *p, long long v) { return p + sprintf(p, "%lld", v); }
static inline char *put_val(char *p, double v) {
    double a = fabs(v);
    if (v == 0.0) {
        *p++ = '0';
        return p;
    }
    if (a < 1e-4) return p + sprintf(p, "%.2E", v);
    return p + sprintf(p, "%.6g", v);
}
static inline char *end_line(char *p, int fields_written, const cfg_t *c) {
    for (int i = fields_written; i < c->width; i++) *p++ = ',';
    if (c->crlf) *p++ = '\r';
    *p++ = '\n';
    return p;
}

/* force pattern ---------------------------------------------------------------------- */
/* Per trecho, 14 equal slots:  rest | single(L) | [both,L,both,R] x2 | both | single(R) | rest | rest
 * => left/right loaded flags per slot.  Chatter: inside long phases, isolated 3-row flips. */
static const int SLOT_L[14] = {0, 1, 1, 1, 1, 0, 1, 1, 1, 0, 1, 0, 0, 0};
static const int SLOT_R[14] = {0, 0, 1, 0, 1, 1, 1, 0, 1, 1, 1, 1, 0, 0};

static inline void plate_state(const cfg_t *c, int64_t row, int *left, int *right) {
    int64_t per_trecho = c->n_dev_rows / 4;
    if (per_trecho < 14 * 16) { /* too short to segment: keep both plates loaded */
        *left = 1;
        *right = 1;
        return;
    }
    int64_t tr = row / per_trecho;
    if (tr > 3) {
        *left = 0;
        *right = 0;
        return;
    }
    int64_t in_tr = row - tr * per_trecho;
    int64_t slot_len = per_trecho / 14;
    int64_t slot = in_tr / slot_len;
    if (slot > 13) slot = 13;
    int64_t in_slot = in_tr - slot * slot_len;
    /* odd trechos walk the other way: swap plates */
    int l = SLOT_L[slot], r = SLOT_R[slot];
    if (tr & 1) {
        int t = l;
        l = r;
        r = t;
    }
    /* chatter: rows [k*97+40, k*97+43) of a slot flip the right plate, only well inside the slot */
    if (slot_len > 400 && in_slot > 50 && in_slot < slot_len - 50) {
        int64_t m = in_slot % 97;
        if (m >= 40 && m < 43 && ((in_slot / 97) % 5 == 0)) r = !r;
    }
    *left = l;
    *right = r;
}

/* rows ------------------------------------------------------------------------------- */
static char *dev_row(char *p, const cfg_t *c, int64_t row) {
    rng_t r = rng_for(c->seed, 1, (uint64_t)row);
    int64_t frame = row / c->subframes + 1, sub = row % c->subframes;
    p = put_int(p, frame);
    *p++ = ',';
    p = put_int(p, sub);
    int fields = 2;
    int loaded[2];
    plate_state(c, row, &loaded[0], &loaded[1]);
    for (int pl = 0; pl < c->n_plates; pl++) {
        int on = pl < 2 ? loaded[pl] : 0;
        double v[9];
        if (on) {
            v[0] = 30.0 * rng_normal(&r);
            v[1] = 30.0 * rng_normal(&r);
            v[2] = -fabs(700.0 + 100.0 * rng_normal(&r)) - 1.0;
            v[3] = 1.0e4 * rng_normal(&r);
            v[4] = 1.0e4 * rng_normal(&r);
            v[5] = 1.0e3 * rng_normal(&r);
            v[6] = 232.0 + 40.0 * rng_normal(&r);
            v[7] = (pl == 0 ? 254.0 : 769.0) + 60.0 * rng_normal(&r);
            v[8] = 0.0;
        } else {
            for (int k = 0; k < 6; k++) v[k] = 0.0;
            v[6] = 232.0;
            v[7] = pl == 0 ? 254.0 : 769.0;
            v[8] = 0.0;
        }
        for (int k = 0; k < 9; k++) {
            *p++ = ',';
            if (!on && k >= 6)
                p = put_int(p, (long long)v[k]);
            else
                p = put_val(p, v[k]);
        }
        fields += 9;
    }
    for (int e = 0; e < c->n_emg; e++) {
        *p++ = ',';
        p = put_val(p, 0.01 * rng_normal(&r));
    }
    fields += c->n_emg;
    return end_line(p, fields, c);
}

static char *traj_row(char *p, const cfg_t *c, int64_t row) {
    rng_t r = rng_for(c->seed, 2, (uint64_t)row);
    p = put_int(p, row + 1);
    *p++ = ',';
    *p++ = '0';
    int fields = 2;
    double t = (double)row / (double)c->f_traj;
    for (int m = 0; m < c->n_markers; m++) {
        double occl = rng_uniform(&r);
        double nx = rng_normal(&r), ny = rng_normal(&r), nz = rng_normal(&r);
        if (occl < c->blank_marker_frac) {
            *p++ = ',';
            *p++ = ',';
            *p++ = ',';
        } else {
            double ph = 0.37 * m;
            double x = 200.0 + 35.0 * m + 300.0 * sin(0.9 * t + ph) + 0.3 * nx;
            double y = 800.0 + 11.0 * m + 700.0 * sin(0.23 * t + 2.0 * ph) + 0.3 * ny;
            double z = 900.0 + 20.0 * (m % 7) + 80.0 * sin(5.1 * t + ph) + 0.3 * nz;
            *p++ = ',';
            p = put_val(p, x);
            *p++ = ',';
            p = put_val(p, y);
            *p++ = ',';
            p = put_val(p, z);
        }
        fields += 3;
    }
    return end_line(p, fields, c);
}

static const char *MUSCLES[16] = {"VL", "RF", "GMED", "TFL", "GMAXS", "GMAXI", "BF", "ST",
                                  "TA", "GASM", "GASL", "SOL", "VM", "ES", "RA", "ADDL"};

static char *header(char *p, const cfg_t *c, int section) {
    char tmp[128];
    if (section == 1) {
        p = put_str(p, "Devices");
        p = end_line(p, 1, c);
        p = put_int(p, c->f_emg);
        p = end_line(p, 1, c);
        /* device names */
        *p++ = ',';
        int fields = 2;
        for (int pl = 0; pl < c->n_plates; pl++) {
            const char *kind[3] = {"Force", "Moment", "CoP"};
            for (int k = 0; k < 3; k++) {
                *p++ = ',';
                sprintf(tmp, "Imported AMTI OR6 Series Force Plate #%d - %s", pl + 1, kind[k]);
                p = put_str(p, tmp);
                *p++ = ',';
                *p++ = ',';
                fields += 3;
            }
        }
        *p++ = ',';
        p = put_str(p, "EMG2000 - Voltage");
        fields += 1;
        p = end_line(p, fields, c);
        /* coordinates */
        p = put_str(p, "Frame,Sub Frame");
        fields = 2;
        for (int pl = 0; pl < c->n_plates; pl++) {
            p = put_str(p, ",Fx,Fy,Fz,Mx,My,Mz,Cx,Cy,Cz");
            fields += 9;
        }
        for (int e = 0; e < c->n_emg; e++) {
            *p++ = ',';
            if (e < 16)
                p = put_str(p, MUSCLES[e]);
            else {
                sprintf(tmp, "M%d", e + 1);
                p = put_str(p, tmp);
            }
            fields++;
        }
        p = end_line(p, fields, c);
        /* units */
        *p++ = ',';
        fields = 2;
        for (int pl = 0; pl < c->n_plates; pl++) {
            p = put_str(p, ",N,N,N,N.mm,N.mm,N.mm,mm,mm,mm");
            fields += 9;
        }
        for (int e = 0; e < c->n_emg; e++) {
            p = put_str(p, ",V");
            fields++;
        }
        p = end_line(p, fields, c);
    } else {
        p = put_str(p, "Trajectories");
        p = end_line(p, 1, c);
        p = put_int(p, c->f_traj);
        p = end_line(p, 1, c);
        *p++ = ',';
        int fields = 2;
        for (int m = 0; m < c->n_markers; m++) {
            *p++ = ',';
            sprintf(tmp, "Subj:MK%02d", m + 1);
            p = put_str(p, tmp);
            *p++ = ',';
            *p++ = ',';
            fields += 3;
        }
        p = end_line(p, fields, c);
        p = put_str(p, "Frame,Sub Frame");
        fields = 2;
        for (int m = 0; m < c->n_markers; m++) {
            p = put_str(p, ",X,Y,Z");
            fields += 3;
        }
        p = end_line(p, fields, c);
        *p++ = ',';
        fields = 2;
        for (int m = 0; m < c->n_markers; m++) {
            p = put_str(p, ",mm,mm,mm");
            fields += 3;
        }
        p = end_line(p, fields, c);
    }
    return p;
}

#define BLOCK_ROWS 2048

/* Generates `n` rows [r0, r0+n) of a section into a fresh malloc'd buffer. */
static char *gen_block(const cfg_t *c, int section, int64_t r0, int64_t n, int64_t *len) {
    size_t per_row = (size_t)c->width * 16 + 64;
    char *buf = (char *)malloc(per_row * (size_t)n);
    if (!buf) return NULL;
    char *p = buf;
    for (int64_t i = 0; i < n; i++) p = section == 1 ? dev_row(p, c, r0 + i) : traj_row(p, c, r0 + i);
    *len = p - buf;
    return buf;
}

/*
 * Returns the number of bytes of the generated file.  If buf is NULL or cap is too small,
 * nothing is written beyond cap and the required size is still returned (call twice).
 */
int64_t ms_synth_vicon(uint8_t *buf, int64_t cap, uint64_t seed, double seconds, int n_plates, int n_emg,
                       int n_markers, int f_emg, int f_traj, int crlf, double blank_marker_frac,
                       int trailing_blank) {
    cfg_t c;
    memset(&c, 0, sizeof(c));
    c.seed = seed;
    c.seconds = seconds;
    c.n_plates = n_plates;
    c.n_emg = n_emg;
    c.n_markers = n_markers;
    c.f_emg = f_emg;
    c.f_traj = f_traj;
    c.crlf = crlf;
    c.blank_marker_frac = blank_marker_frac;
    c.trailing_blank = trailing_blank;
    if (f_traj <= 0 || f_emg <= 0 || f_emg % f_traj != 0 || n_markers < 1 || n_emg < 1 || n_plates < 0) return -1;
    c.subframes = f_emg / f_traj;
    c.n_traj_rows = (int64_t)floor(seconds * f_traj + 0.5);
    if (c.n_traj_rows < 1) c.n_traj_rows = 1;
    c.n_dev_rows = c.n_traj_rows * c.subframes;
    int w1 = 2 + 9 * n_plates + n_emg, w2 = 2 + 3 * n_markers;
    c.width = w1 > w2 ? w1 : w2;

    int64_t nb1 = (c.n_dev_rows + BLOCK_ROWS - 1) / BLOCK_ROWS, nb2 = (c.n_traj_rows + BLOCK_ROWS - 1) / BLOCK_ROWS;
    int64_t nb = nb1 + nb2;
    char **blocks = (char **)calloc((size_t)nb, sizeof(char *));
    int64_t *lens = (int64_t *)calloc((size_t)nb + 1, sizeof(int64_t));
    if (!blocks || !lens) return -2;
    int fail = 0;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t b = 0; b < nb; b++) {
        int section = b < nb1 ? 1 : 2;
        int64_t bi = b < nb1 ? b : b - nb1;
        int64_t total = section == 1 ? c.n_dev_rows : c.n_traj_rows;
        int64_t r0 = bi * BLOCK_ROWS, n = total - r0 < BLOCK_ROWS ? total - r0 : BLOCK_ROWS;
        blocks[b] = gen_block(&c, section, r0, n, &lens[b]);
        if (!blocks[b]) fail = 1;
    }
    if (fail) return -2;

    char hdr1[16384], hdr2[16384], sep[4096];
    int64_t h1 = header(hdr1, &c, 1) - hdr1, h2 = header(hdr2, &c, 2) - hdr2;
    int64_t sl = end_line(sep, 1, &c) - sep;
    int64_t total = h1 + h2 + sl + (trailing_blank ? sl : 0);
    for (int64_t b = 0; b < nb; b++) total += lens[b];

    if (buf && cap >= total) {
        /* offsets */
        int64_t *offs = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nb + 1));
        int64_t pos = 0;
        memcpy(buf + pos, hdr1, (size_t)h1);
        pos += h1;
        for (int64_t b = 0; b < nb1; b++) {
            offs[b] = pos;
            pos += lens[b];
        }
        int64_t sep_pos = pos;
        pos += sl;
        int64_t h2_pos = pos;
        pos += h2;
        for (int64_t b = nb1; b < nb; b++) {
            offs[b] = pos;
            pos += lens[b];
        }
        memcpy(buf + sep_pos, sep, (size_t)sl);
        memcpy(buf + h2_pos, hdr2, (size_t)h2);
        if (trailing_blank) memcpy(buf + pos, sep, (size_t)sl);
#pragma omp parallel for schedule(dynamic, 4)
        for (int64_t b = 0; b < nb; b++) memcpy(buf + offs[b], blocks[b], (size_t)lens[b]);
        free(offs);
    }
    for (int64_t b = 0; b < nb; b++) free(blocks[b]);
    free(blocks);
    free(lens);
    return total;
}
