// better_flow/event_file.h -- event I/O (reference: better_flow_core/include/better_flow/event_file.h,
// I/O parts only: from_file :141-176, to_file_uv :265-289).  All image / colour / arrow rendering of
// the reference class is visualisation and is not part of this tree.
#ifndef BF_EVENT_FILE_H
#define BF_EVENT_FILE_H

#include <charconv>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <system_error>
#include <thread>

#include <better_flow/common.h>
#include <better_flow/event.h>

// Block reader for the "t x y p" text format (addition).  The reference parses with
// `ifstream >> double >> uint >> uint >> bool` (bf_motion_compensator.cpp:190-202, event_file.h:141-176),
// which costs ~1 us per event -- two orders of magnitude more than the minimisation itself on the GPU.
// This reader pulls the file in 4 MiB blocks and converts with std::from_chars, which like the
// strtod behind operator>> is correctly rounded, so every timestamp is the same double.  Same
// termination rule as the reference loop: reading stops at the first record that does not parse
// (p must be 0 or 1, as for operator>>(bool)).
class TextEventReader {
    FILE *fp_;
    bool own_;
    std::vector<char> buf_;
    size_t pos_, end_;
    bool eof_, failed_;

    // make sure a whole line (or the rest of the file) is buffered starting at pos_
    bool fill_line() {
        for (;;) {
            const void *nl = pos_ < end_ ? memchr(buf_.data() + pos_, '\n', end_ - pos_) : nullptr;
            if (nl || eof_) return pos_ < end_;
            if (pos_ > 0) {
                memmove(buf_.data(), buf_.data() + pos_, end_ - pos_);
                end_ -= pos_;
                pos_ = 0;
            }
            if (end_ == buf_.size()) buf_.resize(buf_.size() * 2);   // a line longer than the block
            const size_t got = fread(buf_.data() + end_, 1, buf_.size() - end_, fp_);
            end_ += got;
            if (got == 0) eof_ = true;
        }
    }
    static bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }
    void skip_space() {
        // operator>> skips any whitespace including newlines, so records may span lines
        for (;;) {
            while (pos_ < end_ && is_space(buf_[pos_])) ++pos_;
            if (pos_ < end_ || eof_) return;
            fill_line();
            if (pos_ >= end_) return;
        }
    }
    // token = [pos_, first whitespace); guaranteed fully buffered
    bool token(const char *&b, const char *&e) {
        skip_space();
        if (pos_ >= end_) return false;
        fill_line();
        size_t q = pos_;
        while (q < end_ && !is_space(buf_[q])) ++q;
        b = buf_.data() + pos_;
        e = buf_.data() + q;
        pos_ = q;
        return e > b;
    }
    bool get_double(double &v) {
        const char *b, *e;
        if (!token(b, e)) return false;
        if (*b == '+') ++b;
        const auto r = std::from_chars(b, e, v);
        return r.ec == std::errc() && r.ptr == e;
    }
    bool get_uint(uint &v) {
        const char *b, *e;
        if (!token(b, e)) return false;
        // operator>>(unsigned) follows strtoul: a leading '-' negates the magnitude modulo 2^32
        const bool neg = *b == '-';
        if (*b == '+' || neg) ++b;
        const auto r = std::from_chars(b, e, v);
        if (r.ec != std::errc() || r.ptr != e) return false;
        if (neg) v = 0u - v;
        return true;
    }

    // Fast path for the common shape of a record -- "<digits>[.<digits>] <digits> <digits> <0|1>" on one
    // buffered line: one pass over the characters, no per-token rescans.  The timestamp is converted by
    // Clinger's exact case: a decimal with at most 15 significant digits is an integer N < 2^53 over a power
    // of ten <= 10^22, both exact doubles, and ONE IEEE division N / 10^k is then the correctly rounded value,
    // i.e. the same double strtod / from_chars / operator>> produce (tests/test_reader.py compares them bit
    // for bit).  Anything else (sign, exponent, more digits, a record that spans lines, a malformed token)
    // leaves pos_ untouched and returns false: the general path then decides.
    bool fast_record(double &t, uint &x, uint &y, bool &p) {
        static const double pow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                         1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
        const char *c = buf_.data() + pos_;
        const char *const lim = buf_.data() + end_;
        while (c < lim && (*c == ' ' || *c == '\n' || *c == '\r' || *c == '\t')) ++c;
        const void *nlp = c < lim ? memchr(c, '\n', (size_t)(lim - c)) : nullptr;
        if (!nlp) return false;                              // the line is not completely buffered (or last line: general path)
        const char *const nl = static_cast<const char *>(nlp);
        unsigned long long n = 0;
        int digits = 0, frac = 0;
        while (c < nl && (unsigned)(*c - '0') < 10u) { n = n * 10u + (unsigned)(*c - '0'); ++c; ++digits; }
        if (digits == 0) return false;
        if (c < nl && *c == '.') {
            ++c;
            while (c < nl && (unsigned)(*c - '0') < 10u) { n = n * 10u + (unsigned)(*c - '0'); ++c; ++digits; ++frac; }
        }
        if (digits > 15 || c >= nl || *c != ' ') return false;
        auto field = [&](uint &v) {
            while (c < nl && *c == ' ') ++c;
            unsigned long long u = 0;
            int d = 0;
            while (c < nl && (unsigned)(*c - '0') < 10u) { u = u * 10u + (unsigned)(*c - '0'); ++c; ++d; }
            v = (uint)u;
            return d > 0 && d <= 9;
        };
        uint xv, yv;
        if (!field(xv) || c >= nl || *c != ' ') return false;
        if (!field(yv) || c >= nl || *c != ' ') return false;
        while (c < nl && *c == ' ') ++c;
        if (c >= nl || (*c != '0' && *c != '1')) return false;
        const bool pv = *c == '1';
        ++c;
        while (c < nl && (*c == ' ' || *c == '\r' || *c == '\t')) ++c;
        if (c != nl) return false;
        t = (double)n / pow10[frac];
        x = xv; y = yv; p = pv;
        pos_ = (size_t)(nl - buf_.data());                   // (the newline itself is skipped as whitespace by the next record)
        return true;
    }

public:
    explicit TextEventReader(const std::string &fname) : pos_(0), end_(0), eof_(false), failed_(false) {
        own_ = fname != "-";
        fp_ = own_ ? fopen(fname.c_str(), "rb") : stdin;
        buf_.resize(4u << 20);
        if (!fp_) { eof_ = true; failed_ = true; }
    }
    ~TextEventReader() {
        if (fp_ && own_) fclose(fp_);
    }
    TextEventReader(const TextEventReader &) = delete;
    TextEventReader &operator=(const TextEventReader &) = delete;

    bool opened() const { return fp_ != nullptr; }

    // next "t x y p" record; false at end of input or at the first malformed record
    bool next(double &t, uint &x, uint &y, bool &p) {
        if (failed_) return false;
        if (fast_record(t, x, y, p)) return true;
        uint pv = 0;
        if (!get_double(t) || !get_uint(x) || !get_uint(y) || !get_uint(pv) || pv > 1) {
            failed_ = true;
            return false;
        }
        p = pv != 0;
        return true;
    }
};

// The same records, parsed by a background thread one block ahead of the consumer: text parsing
// (~0.05-0.11 us per event) then overlaps with the minimisation instead of adding to it (the drop-in CLI
// on a 4.5 M-event stream: 0.88 s -> see DESIGN.md).  Record order and values are unchanged.
class PrefetchingTextEventReader {
public:
    struct Rec { double t; uint x, y; bool p; };

    explicit PrefetchingTextEventReader(const std::string &fname, size_t block = 1 << 16)
        : reader_(fname), block_(block), done_(false), stop_(false), pos_(0) {
        worker_ = std::thread([this] { produce(); });
    }
    ~PrefetchingTextEventReader() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_space_.notify_all();
        if (worker_.joinable()) worker_.join();
    }
    PrefetchingTextEventReader(const PrefetchingTextEventReader &) = delete;
    PrefetchingTextEventReader &operator=(const PrefetchingTextEventReader &) = delete;

    bool next(double &t, uint &x, uint &y, bool &p) {
        if (pos_ == cur_.size()) {
            std::unique_lock<std::mutex> lk(m_);
            cv_data_.wait(lk, [this] { return !ready_.empty() || done_; });
            if (ready_.empty()) return false;
            cur_.swap(ready_.front());
            ready_.pop_front();
            pos_ = 0;
            lk.unlock();
            cv_space_.notify_one();
            if (cur_.empty()) return false;
        }
        const Rec &r = cur_[pos_++];
        t = r.t; x = r.x; y = r.y; p = r.p;
        return true;
    }

private:
    void produce() {
        for (;;) {
            std::vector<Rec> blk;
            blk.reserve(block_);
            Rec r;
            while (blk.size() < block_ && reader_.next(r.t, r.x, r.y, r.p)) blk.push_back(r);
            const bool last = blk.size() < block_;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_space_.wait(lk, [this] { return ready_.size() < 4 || stop_; });
                if (stop_) return;
                if (!blk.empty()) ready_.push_back(std::move(blk));
                if (last) done_ = true;
            }
            cv_data_.notify_one();
            if (last) return;
        }
    }

    TextEventReader reader_;
    size_t block_;
    std::thread worker_;
    std::mutex m_;
    std::condition_variable cv_data_, cv_space_;
    std::deque<std::vector<Rec>> ready_;
    bool done_, stop_;
    std::vector<Rec> cur_;
    size_t pos_;
};

class EventFile {
public:
    // "t x y p" text, one event per line; the first timestamp becomes t = 0; note the x/y swap:
    // Event(row = file y, column = file x, t)
    template <class T> static void from_file(T *events, std::string fname) {
        std::cout << "Reading from file... (" << fname << ")" << std::endl << std::flush;
        TextEventReader in(fname);
        ull cnt = 0;
        double t = 0, t_0 = 0;
        uint x = 0, y = 0;
        bool p = false;
        clock_t begin = std::clock();
        if (in.next(t_0, x, y, p)) {
            events->push_back(Event(y, x, FROM_SEC(0)));
            cnt++;
        }
        while (in.next(t, x, y, p)) {
            t -= t_0;
            events->push_back(Event(y, x, FROM_SEC(t)));
            cnt++;
        }
        clock_t end = std::clock();
        if (cnt == 0) {
            std::cout << "Read " << cnt << " events, finished" << std::endl << std::endl << std::flush;
            return;
        }
        std::cout << "Read " << cnt << " events, finished" << std::endl << std::flush;
        std::cout << "Elapsed: " << double(end - begin) / CLOCKS_PER_SEC << " sec." << std::endl << std::flush;
    }

    // Binary input (addition; the text parser costs ~1 us per event): a headerless array of 16-byte
    // records {uint64 t_ns; uint16 x; uint16 y; uint32 polarity}, x = column, y = row; t_ns is used as
    // the event timestamp as is (no re-basing: only differences to the slice start ever matter).
    template <class T> static void from_binary(T *events, std::string fname) {
        std::cout << "Reading from file... (" << fname << ")" << std::endl << std::flush;
        struct Rec { uint64_t t_ns; uint16_t x, y; uint32_t p; };
        static_assert(sizeof(Rec) == 16, "binary event record is 16 bytes");
        std::ifstream in(fname, std::ifstream::in | std::ifstream::binary);
        std::vector<Rec> buf(1 << 16);
        ull cnt = 0;
        while (in) {
            in.read(reinterpret_cast<char *>(buf.data()), (std::streamsize)(buf.size() * sizeof(Rec)));
            const size_t got = size_t(in.gcount()) / sizeof(Rec);
            for (size_t k = 0; k < got; ++k) {
                events->push_back(Event(buf[k].y, buf[k].x, ull(buf[k].t_ns)));
                cnt++;
            }
            if (got < buf.size()) break;
        }
        std::cout << "Read " << cnt << " events, finished" << std::endl << std::flush;
    }

    // The binary format as it lies in the file (16 bytes per event), for callers that build their Events on the fly
    // instead of materialising a cloud of 152-byte records first.
    struct BinaryRecord { uint64_t t_ns; uint16_t x, y; uint32_t p; };
    static std::vector<BinaryRecord> read_binary_records(std::string fname) {
        static_assert(sizeof(BinaryRecord) == 16, "binary event record is 16 bytes");
        std::cout << "Reading from file... (" << fname << ")" << std::endl << std::flush;
        std::vector<BinaryRecord> recs;
        std::ifstream in(fname, std::ifstream::in | std::ifstream::binary | std::ifstream::ate);
        if (!in) return recs;
        const std::streamoff bytes = in.tellg();
        in.seekg(0);
        recs.resize(size_t(bytes) / sizeof(BinaryRecord));
        in.read(reinterpret_cast<char *>(recs.data()), (std::streamsize)(recs.size() * sizeof(BinaryRecord)));
        std::cout << "Read " << recs.size() << " events, finished" << std::endl << std::flush;
        return recs;
    }

    // "t x y 1 v u" per event, fixed 9 decimals; x/y and u/v are swapped back to file convention
    template <class T> static void to_file_uv(T *events, std::string fname) {
        std::cout << "Writing events and flow to file... (" << fname << ")" << std::endl << std::flush;
        std::ofstream out(fname, std::ofstream::out);
        ull cnt = 0;
        clock_t begin = std::clock();
        // std::to_chars(fixed, 9) yields the digits of `ostream << std::fixed << std::setprecision(9)` (both are the
        // correctly rounded decimal expansion; checked on 2 M random doubles incl. -0, nan, inf and 1e300) at a fifth
        // of the cost; the lines are collected in a 1 MB block and written with one call per block.
        std::vector<char> block(1 << 20);
        char *p = block.data(), *const limit = block.data() + block.size() - 1024;   // (a line is < 1024 bytes: 3 doubles <= 330 chars each)
        auto put_f = [&p](double v) { p = std::to_chars(p, p + 340, v, std::chars_format::fixed, 9).ptr; };
        auto put_u = [&p](unsigned long long v) { p = std::to_chars(p, p + 24, v).ptr; };
        for (auto &e : *events) {
            put_f(double(e.timestamp) / 1000000000); *p++ = ' ';
            put_u(e.fr_y); *p++ = ' ';
            put_u(e.fr_x); *p++ = ' ';
            *p++ = '1'; *p++ = ' ';
            put_f(e.best_v); *p++ = ' ';
            put_f(e.best_u); *p++ = '\n';
            if (p > limit) { out.write(block.data(), p - block.data()); p = block.data(); }
            cnt++;
        }
        out.write(block.data(), p - block.data());
        clock_t end = std::clock();
        out.close();
        if (cnt == 0) {
            std::cout << "Written " << cnt << " events, finished" << std::endl << std::endl << std::flush;
            return;
        }
        std::cout << "Written " << cnt << " events, finished" << std::endl << std::flush;
        std::cout << "Elapsed: " << double(end - begin) / CLOCKS_PER_SEC << " sec." << std::endl << std::flush;
    }
};

#endif  // BF_EVENT_FILE_H
