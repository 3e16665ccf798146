// better_flow/accel_lib.h -- the accelerator seam (reference: better_flow_core/include/better_flow/accel_lib.h).
// Same public method names; every one forwards to the CUDA library through the C ABI
// (include/bf_cuda.h).  These are the stage-level calls -- one image, one projection at a time -- kept
// for API parity and for the debug images; OptimizerRolling::run() does NOT loop over them, it hands
// the whole slice to the persistent kernel (bf_minimize / bf_batch_*).
#ifndef BF_ACCEL_LIB_H
#define BF_ACCEL_LIB_H

#include <better_flow/event.h>
#include <better_flow/image.h>
#include <better_flow/object_model.h>
#include <better_flow/opencl_driver.h>

class AccelLib {
public:
    bool gpu_enabled;   // always true here: there is no CPU implementation to fall back to

    AccelLib() : gpu_enabled(true) {}

    // The reference uploads the slice here (accel_lib.h:71-145); buffers are pooled in the context,
    // so this only makes sure the context is large enough.
    template <class T> void init_gpu(T *events, int nRows, int nCols) {
        (void)nRows; (void)nCols;
        CudaDriver::context((long long)events->size(), 1, 1);
    }

    // accel_lib.h:211-217 -> 147-178: mean-timestamp image, (w+scale) x (h+scale)
    template <class T> ImageF get_time_img(T *events, int w, int h, int scale, int x_sh, int y_sh) {
        gather(events);
        ImageF img(w + scale, h + scale);
        bf_ctx *ctx = CudaDriver::context((long long)px_.size(), 1, scale);
        check(bf_time_img(ctx, (int)px_.size(), px_.data(), py_.data(), t_.data(), nz_.data(), w, h, scale, x_sh, y_sh, img.data()),
              "bf_time_img");
        return img;
    }
    template <class T> static ImageF get_time_img_cpu(T *events, int w, int h, int scale, int x_sh, int y_sh) {
        AccelLib a;   // name kept for source compatibility; it runs on the device like everything else
        return a.get_time_img(events, w, h, scale, x_sh, y_sh);
    }

    // accel_lib.h:263-267 -> event.h:99-110
    template <class T> void project_4param_reinit(T *events, double dnx, double dny, double cx, double cy, double div, double crl) {
        gather(events);
        const int n = (int)px_.size();
        if (n == 0) return;
        nx_.resize(n); ny_.resize(n);
        bf_ctx *ctx = CudaDriver::context(n, 1, 1);
        check(bf_project(ctx, n, fx_.data(), fy_.data(), t_.data(), px_.data(), py_.data(), nx_.data(), ny_.data(), dnx, dny, cx, cy,
                         div, crl), "bf_project");
        int i = 0;
        for (auto &e : *events) {
            e.pr_x = px_[i]; e.pr_y = py_[i]; e.nx = nx_[i]; e.ny = ny_[i];
            ++i;
        }
    }

    // accel_lib.h:310-329: nothing is pending -- every call above writes its results back
    template <class T> void writeout_events(T *) {}

    // accel_lib.h:331-398
    void fast_model(ObjectModel &model, ImageF &time_img) { model.update(time_img); }
    ObjectModel fast_model(ImageF &time_img) {
        ObjectModel m;
        m.update(time_img);
        return m;
    }

    // accel_lib.h:400-434 (and Sobel_cpu :513-543): the 3x3 Scharr-weighted gradient images
    void Sobel(ImageF &img, ImageF &grad_x, ImageF &grad_y) {
        grad_x = ImageF(img.rows, img.cols);
        grad_y = ImageF(img.rows, img.cols);
        bf_ctx *ctx = CudaDriver::context(1, 1, 1);
        check(bf_model_from_image(ctx, img.rows, img.cols, img.data(), nullptr, grad_x.data(), grad_y.data()), "bf_model_from_image");
    }
    static void Sobel_cpu(ImageF &img, ImageF &grad_x, ImageF &grad_y) {
        AccelLib a;
        a.Sobel(img, grad_x, grad_y);
    }

private:
    static void check(int rc, const char *what) {
        if (rc < 0) {
            std::cerr << what << " failed: " << bf_last_error() << std::endl;
            std::exit(1);
        }
    }
    template <class T> void gather(T *events) {
        const size_t n = events->size();
        fx_.clear(); fy_.clear(); t_.clear(); px_.clear(); py_.clear(); nz_.clear();
        fx_.reserve(n); fy_.reserve(n); t_.reserve(n); px_.reserve(n); py_.reserve(n); nz_.reserve(n);
        for (auto &e : *events) {
            fx_.push_back((uint16_t)e.fr_x); fy_.push_back((uint16_t)e.fr_y); t_.push_back((int32_t)e.t);
            px_.push_back(e.pr_x); py_.push_back(e.pr_y); nz_.push_back(e.noise ? 1 : 0);
        }
    }
    std::vector<uint16_t> fx_, fy_;
    std::vector<int32_t> t_;
    std::vector<uint8_t> nz_;
    std::vector<double> px_, py_, nx_, ny_;
};

#endif  // BF_ACCEL_LIB_H
