// better_flow/event_file.h -- event I/O (reference: better_flow_core/include/better_flow/event_file.h,
// I/O parts only: from_file :141-176, to_file_uv :265-289).  All image / colour / arrow rendering of
// the reference class is visualisation and is not part of this tree.
#ifndef BF_EVENT_FILE_H
#define BF_EVENT_FILE_H

#include <better_flow/common.h>
#include <better_flow/event.h>

class EventFile {
public:
    // "t x y p" text, one event per line; the first timestamp becomes t = 0; note the x/y swap:
    // Event(row = file y, column = file x, t)
    template <class T> static void from_file(T *events, std::string fname) {
        std::cout << "Reading from file... (" << fname << ")" << std::endl << std::flush;
        std::ifstream in(fname, std::ifstream::in);
        ull cnt = 0;
        double t = 0, t_0 = 0;
        uint x = 0, y = 0;
        bool p = false;
        clock_t begin = std::clock();
        if (in >> t_0 >> x >> y >> p) {
            events->push_back(Event(y, x, FROM_SEC(0)));
            cnt++;
        }
        while (in >> t >> x >> y >> p) {
            t -= t_0;
            events->push_back(Event(y, x, FROM_SEC(t)));
            cnt++;
        }
        clock_t end = std::clock();
        in.close();
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

    // "t x y 1 v u" per event, fixed 9 decimals; x/y and u/v are swapped back to file convention
    template <class T> static void to_file_uv(T *events, std::string fname) {
        std::cout << "Writing events and flow to file... (" << fname << ")" << std::endl << std::flush;
        std::ofstream out(fname, std::ofstream::out);
        ull cnt = 0;
        clock_t begin = std::clock();
        out << std::fixed << std::setprecision(9);
        for (auto &e : *events) {
            out << double(e.timestamp) / 1000000000 << " " << e.fr_y << " " << e.fr_x << " " << 1 << " " << e.best_v << " "
                << e.best_u << "\n";
            cnt++;
        }
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
