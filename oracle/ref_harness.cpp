/*
 * ref_harness.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * A thin extern "C" driver over the UNMODIFIED reference sources, compiled where they lie under
 * /root/reference by oracle/Makefile into oracle/_ref/libracon_ref.so:
 *   racon::createWindow / Window::add_layer / Window::generate_consensus
 *       (/root/reference/src/window.cpp:15-149)
 *   spoa::AlignmentEngine::Create(kNW, m, x, g) + Prealloc(window_length, 5)
 *       (/root/reference/src/polisher.cpp:179-183)
 * It feeds the reference the same flat "window set" arrays that the C-ABI (include/racon_b200.h)
 * and the restated oracle (oracle/poa_oracle.cpp) consume, so all three see identical bytes.
 * Used (a) to pin the restated oracle, (b) as bench.py's cpu_baseline / --impl reference arm.
 */
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "spoa/spoa.hpp"
#include "window.hpp"

namespace {

struct WindowSet {
    uint32_t n_windows;
    const char* bases;
    const char* quals;          // may be null
    const uint64_t* seq_off;    // n_seq + 1
    const uint8_t* seq_has_qual;  // may be null (=> no qualities at all)
    const uint32_t* seq_begin;
    const uint32_t* seq_end;
    const uint32_t* win_first;  // n_windows + 1
    const uint8_t* win_type;    // 0 = kNGS, 1 = kTGS
};

std::shared_ptr<racon::Window> make_window(const WindowSet& ws, uint32_t w, const std::string& dummy) {
    uint32_t s0 = ws.win_first[w];
    uint32_t s1 = ws.win_first[w + 1];
    uint64_t o = ws.seq_off[s0];
    uint32_t bl = static_cast<uint32_t>(ws.seq_off[s0 + 1] - o);
    bool bq = ws.quals != nullptr && ws.seq_has_qual != nullptr && ws.seq_has_qual[s0];
    auto window = racon::createWindow(
        w, w, ws.win_type[w] ? racon::WindowType::kTGS : racon::WindowType::kNGS,
        ws.bases + o, bl, bq ? ws.quals + o : dummy.data(), bl);
    for (uint32_t s = s0 + 1; s < s1; ++s) {
        uint64_t so = ws.seq_off[s];
        uint32_t sl = static_cast<uint32_t>(ws.seq_off[s + 1] - so);
        bool q = ws.quals != nullptr && ws.seq_has_qual != nullptr && ws.seq_has_qual[s];
        window->add_layer(ws.bases + so, sl, q ? ws.quals + so : nullptr, q ? sl : 0,
                          ws.seq_begin[s], ws.seq_end[s]);
    }
    return window;
}

}  // namespace

extern "C" {

/*
 * Runs Window::generate_consensus on every window with `n_threads` host threads (one
 * spoa::AlignmentEngine per thread, as the reference does).  Writes consensus bytes into
 * `out` (window w at out + w*out_stride, length out_len[w]) and polished[w] = return value.
 * Returns the wall-clock seconds spent in the consensus loop only (steady_clock), or <0 on error.
 */
double ref_poa_consensus(uint32_t n_windows, const char* bases, const char* quals,
                         const uint64_t* seq_off, const uint8_t* seq_has_qual,
                         const uint32_t* seq_begin, const uint32_t* seq_end,
                         const uint32_t* win_first, const uint8_t* win_type, int8_t match,
                         int8_t mismatch, int8_t gap, uint32_t window_length, int trim,
                         uint32_t n_threads, char* out, uint32_t out_stride, uint32_t* out_len,
                         uint8_t* polished) {
    WindowSet ws{n_windows, bases, quals, seq_off, seq_has_qual, seq_begin, seq_end, win_first, win_type};
    uint32_t max_bl = 1;
    for (uint32_t w = 0; w < n_windows; ++w) {
        uint32_t s0 = win_first[w];
        uint32_t bl = static_cast<uint32_t>(seq_off[s0 + 1] - seq_off[s0]);
        if (bl > max_bl) max_bl = bl;
    }
    std::string dummy(max_bl, '!');
    std::vector<std::shared_ptr<racon::Window>> windows;
    windows.reserve(n_windows);
    for (uint32_t w = 0; w < n_windows; ++w) windows.push_back(make_window(ws, w, dummy));

    if (n_threads == 0) n_threads = 1;
    std::vector<std::shared_ptr<spoa::AlignmentEngine>> engines;
    for (uint32_t t = 0; t < n_threads; ++t) {
        engines.emplace_back(spoa::AlignmentEngine::Create(spoa::AlignmentType::kNW, match, mismatch, gap));
        engines.back()->Prealloc(window_length, 5);
    }

    std::atomic<uint32_t> next(0);
    std::vector<uint8_t> flags(n_windows, 0);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](uint32_t tid) {
        for (;;) {
            uint32_t w = next.fetch_add(1);
            if (w >= n_windows) break;
            flags[w] = windows[w]->generate_consensus(engines[tid], trim != 0) ? 1 : 0;
        }
    };
    if (n_threads == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (uint32_t t = 0; t < n_threads; ++t) pool.emplace_back(work, t);
        for (auto& th : pool) th.join();
    }
    auto t1 = std::chrono::steady_clock::now();

    for (uint32_t w = 0; w < n_windows; ++w) {
        const std::string& c = windows[w]->consensus();
        if (c.size() > out_stride) return -1.0;
        std::memcpy(out + static_cast<uint64_t>(w) * out_stride, c.data(), c.size());
        out_len[w] = static_cast<uint32_t>(c.size());
        polished[w] = flags[w];
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
