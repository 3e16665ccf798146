#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>
extern "C" int st_stream(int config, int n, const uint32_t *fr_x, const uint32_t *fr_y, const uint64_t *ts,
                         unsigned long long ev_refresh, unsigned long long time_refresh_ns, int scale, int max_iter, int stm_disable,
                         int flush, int batch, int local, int lazy, int max_slices, double *models, long long *info, double *uv);
int main() {
    const int n = 160000;
    std::mt19937 g(7);
    std::vector<uint32_t> fx(n), fy(n); std::vector<uint64_t> ts(n);
    uint64_t t = 1000;
    for (int i = 0; i < n; i++) { fx[i] = 20 + g() % 140; fy[i] = 20 + g() % 200; t += 500 + g() % 1500; ts[i] = t; }
    std::vector<double> models(512 * 11), uv(512 * 14); std::vector<long long> info(512 * 3);
    for (int lazy : {1, 2, 3, 4})
        for (int stm : {0, 1})
            for (int cfg : {0, 1}) {
                int k = st_stream(cfg, n, fx.data(), fy.data(), ts.data(), lazy == 2 ? 40000 : 20000, 33000000ull, 3, 2, stm, 1, 1, 0, lazy, 512, models.data(), info.data(), uv.data());
                printf("lazy %d stm %d cfg %d -> %d slices, last total_dx %.6g\n", lazy, stm, cfg, k, models[11 * (k - 1) + 7]);
            }
    return 0;
}
