// Host build of the device field parser (muscle_synergies_b200/csrc/ms_parse_double.cuh)
// for CPU-side fuzzing against CPython float() / glibc strtod.  TEST INFRASTRUCTURE ONLY:
// nothing in the product links this.
//   g++ -O2 -shared -fPIC tests/native/parse_harness.cpp -o tests/native/libms_parse_harness.so
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../muscle_synergies_b200/csrc/ms_parse_double.cuh"

extern "C" {

int ms_host_parse(const char* s, int len, uint64_t* bits) {
    return ms_parse_field((const uint8_t*)s, (const uint8_t*)s + len, bits);
}

// which tier would handle it: forces tier 2 / tier 3 for direct testing
uint64_t ms_host_eisel_lemire(uint64_t w, int64_t q) { return ms_eisel_lemire(w, q); }
uint64_t ms_host_exact(const char* s, int mant_len, int64_t exp10) {
    return ms_exact_decimal((const uint8_t*)s, (const uint8_t*)s + mant_len, exp10);
}

// Parses `n` NUL-separated strings packed in `blob`; writes bits and status.
void ms_host_parse_many(const char* blob, const int64_t* offsets, int64_t n, uint64_t* bits, int32_t* status) {
    for (int64_t i = 0; i < n; i++) {
        uint64_t b = 0;
        status[i] = ms_parse_field((const uint8_t*)blob + offsets[i], (const uint8_t*)blob + offsets[i + 1], &b);
        bits[i] = b;
    }
}

static uint64_t sm64(uint64_t* x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// Fuzz against glibc strtod (correctly rounded).  mode 0: random bit patterns printed with
// %.17g; 1: printed with 1..25 significant digits (%.*e); 2: random digit strings of 1..40
// digits with a random decimal point and exponent in [-345, 310]; 3: Vicon-like short fields.
// Forces tier 3 as well when force_exact != 0 (compares ms_exact_decimal against strtod).
int64_t ms_host_fuzz(uint64_t seed, int64_t n, int mode, int force_exact, char* first_bad, int first_bad_cap) {
    int64_t bad = 0;
    char buf[256];
    uint64_t st = seed;
    for (int64_t i = 0; i < n; i++) {
        int len = 0;
        if (mode == 0 || mode == 1) {
            uint64_t b = sm64(&st);
            if (((b >> 52) & 0x7FF) == 0x7FF) b &= ~(1ull << 62);  // avoid inf/nan
            double d;
            memcpy(&d, &b, 8);
            if (mode == 0)
                len = snprintf(buf, sizeof buf, "%.17g", d);
            else
                len = snprintf(buf, sizeof buf, "%.*e", (int)(sm64(&st) % 25), d);
        } else if (mode == 2) {
            int nd = 1 + (int)(sm64(&st) % 40);
            int dot = (int)(sm64(&st) % (nd + 1));
            char* p = buf;
            if (sm64(&st) & 1) *p++ = '-';
            for (int k = 0; k < nd; k++) {
                if (k == dot) *p++ = '.';
                *p++ = (char)('0' + sm64(&st) % 10);
            }
            if (dot == nd && (sm64(&st) & 1)) *p++ = '.';
            if (sm64(&st) % 4) {
                int e = (int)(sm64(&st) % 656) - 345;
                p += sprintf(p, "%c%d", (sm64(&st) & 1) ? 'e' : 'E', e);
            }
            *p = 0;
            len = (int)(p - buf);
        } else {
            double v;
            uint64_t r = sm64(&st);
            double u = (double)(sm64(&st) >> 11) / 9007199254740992.0 - 0.5;
            switch (r % 4) {
                case 0: v = u * 0.05; break;
                case 1: v = u * 3000.0; break;
                case 2: v = u * 1e5; break;
                default: v = u * 2e-4; break;
            }
            if (v != 0 && (v < 0 ? -v : v) < 1e-4)
                len = snprintf(buf, sizeof buf, "%.2E", v);
            else
                len = snprintf(buf, sizeof buf, "%.6g", v);
        }
        double ref = strtod(buf, NULL);
        uint64_t rb;
        memcpy(&rb, &ref, 8);
        uint64_t got = 0;
        int stt;
        if (force_exact) {
            // split mantissa / exponent by hand for the exact routine
            const char* s = buf;
            uint64_t sign = 0;
            if (*s == '-') { sign = 1ull << 63; s++; }
            const char* m = s;
            while (*m && *m != 'e' && *m != 'E') m++;
            long long e10 = *m ? atoll(m + 1) : 0;
            got = sign | ms_exact_decimal((const uint8_t*)s, (const uint8_t*)m, e10);
            stt = 0;
        } else {
            stt = ms_parse_field((const uint8_t*)buf, (const uint8_t*)buf + len, &got);
        }
        if (stt != 0 || got != rb) {
            if (bad == 0 && first_bad) snprintf(first_bad, first_bad_cap, "%s -> got %016llx want %016llx st %d", buf,
                                                (unsigned long long)got, (unsigned long long)rb, stt);
            bad++;
        }
    }
    return bad;
}
}
