/*
 * C restatement of the reference loader's data path.  TEST INFRASTRUCTURE ONLY - nothing in
 * the product links or calls this (only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs may).
 *
 * It follows, row by row, what the reference does in Python
 * (elvis-sik/muscle_synergies, src/muscle_synergies/vicon_data/):
 *   load_csv.py:21-31     text-mode universal newlines + csv.reader (excel dialect; quoted
 *                         fields are NOT handled here: returns MSO_E_QUOTE)
 *   reader.py:116-130     strip every field / drop trailing empties (blank-row test :886-901)
 *   reader.py:760-794     coordinates line: number of fields after trimming = num_cols
 *   reader.py:927-948     first num_cols fields: "" -> None (NaN), else float()
 *   aggregator.py:96-124  columns 0,1 (Frame, Sub Frame) are kept by no device
 * float() is restated with glibc strtod (correctly rounded, like CPython's dtoa) behind a
 * grammar check that accepts exactly what float() accepts for ASCII text.
 *
 * It exists because the Python oracle (oracle/vicon_oracle.py, the algorithm-faithful
 * "port") needs minutes on the 500 MB configuration; this one checks the CUDA output
 * bit for bit at full size in seconds.  tests/test_oracle_golden.py pins it against the
 * Python oracle and the reference-generated golden vectors.
 *
 * Build: gcc -O2 -shared -fPIC oracle/vicon_oracle_c.c -o oracle/libvicon_oracle.so
 */
#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MSO_OK 0
#define MSO_E_QUOTE (-1)   /* quoted field: not restated here */
#define MSO_E_HEADER (-2)  /* structure this restatement does not handle (use the Python oracle) */
#define MSO_E_FLOAT (-3)   /* a field float() rejects; *err_line holds the 1-based line */

typedef struct {
    int64_t n_rows[2];   /* data rows of Devices / Trajectories */
    int32_t num_cols[2]; /* parsed fields per row */
    int64_t err_line;
} mso_info;

static int is_strip_space(unsigned c) { return c == 32 || (c >= 9 && c <= 13) || (c >= 28 && c <= 31); }
static int is_float_space(unsigned c) { return c == 32 || (c >= 9 && c <= 13); }

/* one physical line: [p, e) content, next = start of the following line */
static const char *line_end(const char *p, const char *end, const char **next) {
    const char *q = p;
    while (q < end && *q != '\n' && *q != '\r') q++;
    if (q == end)
        *next = end;
    else if (*q == '\r' && q + 1 < end && q[1] == '\n')
        *next = q + 2;
    else
        *next = q + 1;
    return q;
}

static int line_is_blank(const char *p, const char *e) {
    for (; p < e; p++)
        if (*p != ',' && !is_strip_space((unsigned char)*p)) return 0;
    return 1;
}

/* number of fields after stripping each and dropping trailing empty ones */
static int trimmed_field_count(const char *p, const char *e) {
    int count = 0, idx = 0;
    const char *f = p;
    for (const char *q = p;; q++) {
        if (q == e || *q == ',') {
            const char *a = f, *b = q;
            while (a < b && is_strip_space((unsigned char)*a)) a++;
            idx++;
            if (a < b) count = idx;
            if (q == e) break;
            f = q + 1;
        }
    }
    return count;
}

/* float(field) for a non-empty ASCII field; returns 0 on success */
static int py_float(const char *s, const char *e, double *out) {
    char tmp[512];
    while (s < e && is_float_space((unsigned char)*s)) s++;
    while (e > s && is_float_space((unsigned char)e[-1])) e--;
    if (s == e || e - s >= (long)sizeof tmp) return -1;
    /* underscores only between digits; everything else checked against the grammar */
    int n = 0;
    char prev = 0;
    for (const char *p = s; p < e; p++) {
        char c = *p;
        if ((unsigned char)c >= 0x80 || c == 0) return -1;
        if (c == '_') {
            if (!(prev >= '0' && prev <= '9')) return -1;
            if (p + 1 >= e || !(p[1] >= '0' && p[1] <= '9')) return -1;
        } else {
            tmp[n++] = c;
        }
        prev = c;
    }
    tmp[n] = 0;
    const char *p = tmp;
    if (*p == '+' || *p == '-') p++;
    char low[16];
    int ln = 0;
    for (const char *q = p; *q && ln < 15; q++) low[ln++] = (char)tolower((unsigned char)*q);
    low[ln] = 0;
    if (!strcmp(low, "inf") || !strcmp(low, "infinity")) {
        *out = tmp[0] == '-' ? -INFINITY : INFINITY;
        return 0;
    }
    if (!strcmp(low, "nan")) {
        union {
            uint64_t u;
            double d;
        } v;
        v.u = tmp[0] == '-' ? 0xfff8000000000000ull : 0x7ff8000000000000ull;
        *out = v.d;
        return 0;
    }
    /* digits [. digits] [e[+-]digits] with at least one mantissa digit */
    int nd = 0;
    while (*p >= '0' && *p <= '9') p++, nd++;
    if (*p == '.') {
        p++;
        while (*p >= '0' && *p <= '9') p++, nd++;
    }
    if (nd == 0) return -1;
    if (*p == 'e' || *p == 'E') {
        p++;
        if (*p == '+' || *p == '-') p++;
        if (!(*p >= '0' && *p <= '9')) return -1;
        while (*p >= '0' && *p <= '9') p++;
    }
    if (*p) return -1;
    *out = strtod(tmp, NULL);
    return 0;
}

/* Walks the file.  With out == NULL only counts; otherwise fills out[s] as a ROW-major
 * (n_rows[s], num_cols[s] - 2) float64 array (NaN for empty or missing cells). */
static int walk(const char *buf, int64_t n, mso_info *info, double *out0, double *out1) {
    const char *p = buf, *end = buf + n;
    int section = 0, state = 1; /* states 1..5 header lines, 6 data */
    int64_t line = 0;
    int num_cols = 0;
    double *outs[2] = {out0, out1};
    union {
        uint64_t u;
        double d;
    } nanv;
    nanv.u = 0x7ff8000000000000ull;
    info->n_rows[0] = info->n_rows[1] = 0;
    info->num_cols[0] = info->num_cols[1] = 0;
    info->err_line = 0;
    if (memchr(buf, '"', (size_t)n)) return MSO_E_QUOTE;
    while (p < end) {
        const char *next;
        const char *e = line_end(p, end, &next);
        line++;
        if (state == 6) {
            if (line_is_blank(p, e)) {
                section++;
                state = 1;
                if (section > 1) {
                    if (next < end) {
                        info->err_line = line + 1;
                        return MSO_E_HEADER;
                    }
                    break;
                }
            } else {
                int64_t r = info->n_rows[section]++;
                double *row = outs[section] ? outs[section] + r * (int64_t)(num_cols - 2) : NULL;
                int col = 0;
                const char *f = p;
                for (const char *q = p; row && col < num_cols; q++) {
                    if (q == e || *q == ',') {
                        double v = nanv.d;
                        if (q > f && py_float(f, q, &v) != 0) {
                            info->err_line = line;
                            return MSO_E_FLOAT;
                        }
                        if (row && col >= 2) row[col - 2] = v;
                        col++;
                        if (q == e) break;
                        f = q + 1;
                    }
                }
                if (row)
                    for (int c = col < 2 ? 2 : col; c < num_cols; c++) row[c - 2] = nanv.d;
            }
        } else {
            if (state == 1) {
                const char *word = section == 0 ? "Devices" : "Trajectories";
                size_t wl = strlen(word);
                if ((size_t)(e - p) < wl || memcmp(p, word, wl) != 0) {
                    info->err_line = line;
                    return MSO_E_HEADER;
                }
            } else if (state == 4) {
                num_cols = trimmed_field_count(p, e);
                if (num_cols < 3) {
                    info->err_line = line;
                    return MSO_E_HEADER;
                }
                info->num_cols[section] = num_cols;
            }
            state++;
        }
        p = next;
    }
    return MSO_OK;
}

int mso_count(const char *buf, int64_t n, mso_info *info) { return walk(buf, n, info, NULL, NULL); }

int mso_parse(const char *buf, int64_t n, mso_info *info, double *out_devices, double *out_traj) {
    return walk(buf, n, info, out_devices, out_traj);
}
