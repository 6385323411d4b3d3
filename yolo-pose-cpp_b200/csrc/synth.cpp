// synth.cpp — synthetic YOLO-Pose head tensors for tests and benchmarks (host code).
//
// There are no model weights or videos in this environment, so the input to the path is
// generated: temporally coherent skeletons (the template of the reference's
// benchmark.cpp:32-42) moving on closed periodic curves, K duplicate firing anchors per
// person (what a real head emits around each person), confidence flicker and optional
// occlusion gaps, on top of a low-confidence random background.  Values come from a
// counter-based generator (splitmix64 of seed/stream/frame/entity/field) and plain
// arithmetic only (no libm), so any stream-frame can be generated independently, in any
// order, on any thread, with identical bytes.
//
// Layout produced: one [56, N] fp32 slab per stream-frame, channel-major exactly as the
// TensorRT engine emits it (reference gpu_postprocess.cu:44-47): rows 0-3 cx,cy,w,h,
// row 4 confidence, rows 5+3k..7+3k keypoint k (x, y, conf).
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

extern "C" {

typedef struct pb_synth_config {
    int num_anchors;        // 8400 (640x640) or 33600 (1280x1280): (S/8)^2 + (S/16)^2 + (S/32)^2
    int canvas;             // S
    int persons;            // P per stream
    int period;             // motion period in frames; frame f and f+period are identical
    int clumps;             // 0: centres uniform; >0: centres drawn around this many clump centres
    int occlusion;          // 1: every person disappears for two intervals per period
    float kp_drop_prob;     // probability that a keypoint's confidence drops below 0.15
    float max_speed;        // px/frame cap on centre motion
    unsigned long long seed;
} pb_synth_config;

}  // extern "C"

namespace {

inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
inline uint64_t key(uint64_t seed, uint64_t stream, uint64_t frame, uint64_t entity, uint64_t field) {
    uint64_t h = mix64(seed ^ 0x5EEDB200ULL);
    h = mix64(h ^ (stream * 0xD6E8FEB86659FD93ULL));
    h = mix64(h ^ (frame * 0xA0761D6478BD642FULL));
    h = mix64(h ^ (entity * 0xE7037ED1A0B428DBULL));
    return mix64(h ^ (field * 0x8EBC6AF09C88C6E3ULL));
}
inline float u01(uint64_t h) { return (float)(h >> 40) * (1.0f / 16777216.0f); }
inline float uni(uint64_t h, float lo, float hi) { return lo + (hi - lo) * u01(h); }
// ~N(0,1): sum of four uniforms (Irwin-Hall), no libm.
inline float gauss(uint64_t h) {
    float s = 0.0f;
    for (int i = 0; i < 4; ++i) { s += u01(h); h = mix64(h); }
    return (s - 2.0f) * 1.7320508f;
}
// C1-continuous periodic wave with period 1 and range [-1,1] (parabolic sine).
inline float wave(float t) {
    t -= (float)(int)t;
    if (t < 0) t += 1.0f;
    float x = 2.0f * t - 1.0f;
    float ax = x < 0 ? -x : x;
    return -4.0f * x * (1.0f - ax);
}

const float kTemplate[17][2] = {
    {0.0f, -1.5f}, {-0.1f, -1.6f}, {0.1f, -1.6f}, {-0.2f, -1.5f}, {0.2f, -1.5f},
    {-0.5f, -1.0f}, {0.5f, -1.0f}, {-0.8f, -0.3f}, {0.8f, -0.3f}, {-1.0f, 0.3f}, {1.0f, 0.3f},
    {-0.3f, 0.0f}, {0.3f, 0.0f}, {-0.3f, 0.8f}, {0.3f, 0.8f}, {-0.3f, 1.5f}, {0.3f, 1.5f}};

struct Person {
    bool visible;
    float kp[17][3];
    float box[4];   // cx cy w h
    float cx, cy, s;
};

// True (noise-free) state of person p of `stream` at `frame`.
Person person_at(const pb_synth_config& c, int stream, int frame, int p) {
    Person P;
    const float S = (float)c.canvas;
    const int F = c.period > 0 ? c.period : 1;
    const int f = ((frame % F) + F) % F;
    const uint64_t sd = c.seed;
    float bx, by;
    if (c.clumps > 0) {
        int cl = (int)(key(sd, stream, 0, p, 1) % (uint64_t)c.clumps);
        float ccx = uni(key(sd, stream, 0, 100000 + cl, 2), 0.15f * S, 0.85f * S);
        float ccy = uni(key(sd, stream, 0, 100000 + cl, 3), 0.15f * S, 0.85f * S);
        bx = ccx + 60.0f * gauss(key(sd, stream, 0, p, 4));
        by = ccy + 60.0f * gauss(key(sd, stream, 0, p, 5));
    } else {
        bx = uni(key(sd, stream, 0, p, 2), 0.1f * S, 0.9f * S);
        by = uni(key(sd, stream, 0, p, 3), 0.15f * S, 0.85f * S);
    }
    float s = (c.canvas >= 1280) ? uni(key(sd, stream, 0, p, 6), 30.0f, 70.0f)
                                 : uni(key(sd, stream, 0, p, 6), 25.0f, 60.0f);
    int kx = 1 + (int)(key(sd, stream, 0, p, 7) & 1), ky = 1 + (int)(key(sd, stream, 0, p, 8) & 1);
    // amplitude limited so that |v| <= max_speed: v_max = A * 8 * k / F for this wave.
    float amax_x = c.max_speed * (float)F / (8.0f * kx), amax_y = c.max_speed * (float)F / (8.0f * ky);
    float ax = uni(key(sd, stream, 0, p, 9), 5.0f, 40.0f), ay = uni(key(sd, stream, 0, p, 10), 3.0f, 25.0f);
    if (ax > amax_x) ax = amax_x;
    if (ay > amax_y) ay = amax_y;
    float phx = u01(key(sd, stream, 0, p, 11)), phy = u01(key(sd, stream, 0, p, 12));
    float t = (float)f / (float)F;
    P.cx = bx + ax * wave(kx * t + phx);
    P.cy = by + ay * wave(ky * t + phy);
    P.s = s;
    P.visible = true;
    if (c.occlusion) {
        for (int g = 0; g < 2; ++g) {
            int start = (int)(key(sd, stream, 0, p, 20 + g) % (uint64_t)F);
            int len = 5 + (int)(key(sd, stream, 0, p, 30 + g) % 41);   // 5..45
            if (len > F / 3) len = F / 3;
            int rel = ((f - start) % F + F) % F;
            if (rel < len) P.visible = false;
        }
    }
    float lx = 1e9f, ly = 1e9f, hx = -1e9f, hy = -1e9f;
    for (int k = 0; k < 17; ++k) {
        float wig = 0.05f * s * wave((1 + k % 3) * t + u01(key(sd, stream, 0, p, 40 + k)));
        float x = P.cx + kTemplate[k][0] * s + wig;
        float y = P.cy + kTemplate[k][1] * s + 0.5f * wig;
        float conf = uni(key(sd, stream, f + 1, p, 60 + k), 0.3f, 1.0f);
        if (u01(key(sd, stream, f + 1, p, 80 + k)) < c.kp_drop_prob)
            conf = uni(key(sd, stream, f + 1, p, 100 + k), 0.0f, 0.15f);
        P.kp[k][0] = x; P.kp[k][1] = y; P.kp[k][2] = conf;
        if (x < lx) lx = x; if (x > hx) hx = x;
        if (y < ly) ly = y; if (y > hy) hy = y;
    }
    float w = (hx - lx) * 1.1f, h = (hy - ly) * 1.1f;
    P.box[0] = (lx + hx) * 0.5f; P.box[1] = (ly + hy) * 0.5f; P.box[2] = w; P.box[3] = h;
    return P;
}

void fill_background(const pb_synth_config& c, int stream, int frame, float* out) {
    const int N = c.num_anchors;
    const float S = (float)c.canvas;
    const int F = c.period > 0 ? c.period : 1;
    const int f = ((frame % F) + F) % F;
    const uint64_t base = key(c.seed, stream, f + 1, 0xFFFFFF, 0);
    const size_t total = (size_t)56 * N;
    for (int r = 0; r < 56; ++r) {
        const float scale = (r == 4) ? 0.2f : S;
        float* row = out + (size_t)r * N;
        size_t e0 = (size_t)r * N;
        for (int a = 0; a < N; a += 2) {
            uint64_t h = mix64(base + (e0 + a) / 2 * 0x2545F4914F6CDD1DULL);
            row[a] = (float)((h >> 8) & 0xFFFFFF) * (1.0f / 16777216.0f) * scale;
            if (a + 1 < N) row[a + 1] = (float)(h >> 40) * (1.0f / 16777216.0f) * scale;
        }
    }
    (void)total;
}

void gen_head(const pb_synth_config& c, int stream, int frame, float* out) {
    fill_background(c, stream, frame, out);
    const int N = c.num_anchors;
    const int S = c.canvas;
    const int F = c.period > 0 ? c.period : 1;
    const int f = ((frame % F) + F) % F;
    const int g[3] = {S / 8, S / 16, S / 32};
    const int stride[3] = {8, 16, 32};
    const int base[3] = {0, g[0] * g[0], g[0] * g[0] + g[1] * g[1]};
    for (int p = 0; p < c.persons; ++p) {
        Person P = person_at(c, stream, frame, p);
        if (!P.visible) continue;
        int K = 5 + (int)(key(c.seed, stream, f + 1, p, 200) % 5);          // 5..9 firing anchors
        float tx = P.cx, ty = P.cy - 0.5f * P.s;                            // torso point
        for (int j = 0; j < K; ++j) {
            int lv = j % 3;
            int ox = j / 3 - 1;
            int oy = (int)(key(c.seed, stream, f + 1, p, 210 + j) % 3) - 1;
            int gx = (int)(tx / stride[lv]) + ox, gy = (int)(ty / stride[lv]) + oy;
            if (gx < 0) gx = 0; if (gx >= g[lv]) gx = g[lv] - 1;
            if (gy < 0) gy = 0; if (gy >= g[lv]) gy = g[lv] - 1;
            int a = base[lv] + gy * g[lv] + gx;
            if (a >= N) continue;
            uint64_t e = (uint64_t)p * 16 + j;
            out[(size_t)4 * N + a] = uni(key(c.seed, stream, f + 1, e, 300), 0.35f, 0.95f);
            for (int r = 0; r < 4; ++r)
                out[(size_t)r * N + a] = P.box[r] + 1.5f * gauss(key(c.seed, stream, f + 1, e, 310 + r));
            for (int k = 0; k < 17; ++k) {
                out[(size_t)(5 + 3 * k) * N + a] = P.kp[k][0] + 1.5f * gauss(key(c.seed, stream, f + 1, e, 320 + 2 * k));
                out[(size_t)(6 + 3 * k) * N + a] = P.kp[k][1] + 1.5f * gauss(key(c.seed, stream, f + 1, e, 321 + 2 * k));
                float cf = P.kp[k][2] + 0.02f * gauss(key(c.seed, stream, f + 1, e, 400 + k));
                if (cf < 0.0f) cf = 0.0f; if (cf > 1.0f) cf = 1.0f;
                out[(size_t)(7 + 3 * k) * N + a] = cf;
            }
        }
    }
}

}  // namespace

extern "C" {

// One [56,N] slab.
void pb_synth_head(const pb_synth_config* c, int stream, int frame, float* out) {
    gen_head(*c, stream, frame, out);
}

// nstreams x nframes slabs.  frame_major = 1: out[f][b][56][N] (one batch per frame,
// the layout a batched engine emits); 0: out[b][f][56][N].
void pb_synth_heads(const pb_synth_config* c, int stream0, int nstreams, int frame0, int nframes,
                    int frame_major, float* out, int n_threads) {
    const size_t slab = (size_t)56 * c->num_anchors;
    const int total = nstreams * nframes;
    if (n_threads < 1) n_threads = 1;
    auto work = [&](int tid) {
        for (int i = tid; i < total; i += n_threads) {
            int b = i / nframes, f = i % nframes;
            size_t off = frame_major ? ((size_t)f * nstreams + b) : ((size_t)b * nframes + f);
            gen_head(*c, stream0 + b, frame0 + f, out + off * slab);
        }
    };
    if (n_threads == 1) { work(0); return; }
    std::vector<std::thread> th;
    for (int i = 0; i < n_threads; ++i) th.emplace_back(work, i);
    for (auto& t : th) t.join();
}

// Direct tracker input (config 5: no head): one detection per visible person, poses with
// measurement noise, scores in (0.35,0.95) sorted descending like a post-NMS list.
// Returns the number of detections written (<= persons).
int pb_synth_dets(const pb_synth_config* c, int stream, int frame, float* poses, float* scores) {
    const int F = c->period > 0 ? c->period : 1;
    const int f = ((frame % F) + F) % F;
    struct Item { float score; int p; };
    std::vector<Item> items;
    for (int p = 0; p < c->persons; ++p) {
        Person P = person_at(*c, stream, frame, p);
        if (!P.visible) continue;
        items.push_back({uni(key(c->seed, stream, f + 1, p, 500), 0.35f, 0.95f), p});
    }
    // insertion sort, descending, stable
    for (size_t i = 1; i < items.size(); ++i) {
        Item it = items[i];
        size_t j = i;
        while (j > 0 && items[j - 1].score < it.score) { items[j] = items[j - 1]; --j; }
        items[j] = it;
    }
    int n = 0;
    for (const Item& it : items) {
        Person P = person_at(*c, stream, frame, it.p);
        for (int k = 0; k < 17; ++k) {
            poses[n * 51 + k * 3 + 0] = P.kp[k][0] + 1.5f * gauss(key(c->seed, stream, f + 1, it.p, 520 + 2 * k));
            poses[n * 51 + k * 3 + 1] = P.kp[k][1] + 1.5f * gauss(key(c->seed, stream, f + 1, it.p, 521 + 2 * k));
            poses[n * 51 + k * 3 + 2] = P.kp[k][2];
        }
        scores[n] = it.score;
        ++n;
    }
    return n;
}

}  // extern "C"
