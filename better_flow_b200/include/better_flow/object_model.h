// better_flow/object_model.h -- the 4-parameter motion model (reference:
// better_flow_core/include/better_flow/object_model.h, src/object_model.cpp).  Same public fields.
// The image reductions (center_of_mass / compute) run on the GPU: update(img) forwards to the CUDA
// library; there is no host implementation of them in this tree.
#ifndef BF_OBJECT_MODEL_H
#define BF_OBJECT_MODEL_H

#include <better_flow/common.h>
#include <better_flow/image.h>
#include <better_flow/opencl_driver.h>

class ObjectModel {
public:
    double cx, cy, dx, dy, rot, div;
    uint cnt;
    double total_dx, total_dy, total_rot, total_div;

    ObjectModel() : cx(0), cy(0), dx(0), dy(0), rot(0), div(0), cnt(0), total_dx(0), total_dy(0), total_rot(0), total_div(0) {}
    explicit ObjectModel(ImageF &time_img) : ObjectModel() { update(time_img); }
    explicit ObjectModel(const bf_model &m) { from_pod(m); }

    // center_of_mass + compute on the device (object_model.h:31-34 -> object_model.cpp:103-126, 4-39)
    void update(ImageF &time_img) {
        double out7[7];
        bf_ctx *ctx = CudaDriver::context(1, 1, 1);
        if (bf_model_from_image(ctx, time_img.rows, time_img.cols, time_img.data(), out7, nullptr, nullptr) != BF_OK) {
            std::cerr << "bf_model_from_image: " << bf_last_error() << std::endl;
            std::exit(1);
        }
        cx = out7[0]; cy = out7[1]; dx = out7[2]; dy = out7[3]; rot = out7[4]; div = out7[5]; cnt = uint(out7[6]);
    }

    // object_model.h:48-53
    void update_accumulators(float d1, float d2, float d3, float d4) {
        total_rot += rot / d1;
        total_div += div / d2;
        total_dx += dx / d3;
        total_dy += dy / d4;
    }

    bf_model to_pod() const {
        bf_model m;
        m.cx = cx; m.cy = cy; m.dx = dx; m.dy = dy; m.rot = rot; m.div = div; m.cnt = cnt; m.pad_ = 0;
        m.total_dx = total_dx; m.total_dy = total_dy; m.total_rot = total_rot; m.total_div = total_div;
        return m;
    }
    void from_pod(const bf_model &m) {
        cx = m.cx; cy = m.cy; dx = m.dx; dy = m.dy; rot = m.rot; div = m.div; cnt = m.cnt;
        total_dx = m.total_dx; total_dy = m.total_dy; total_rot = m.total_rot; total_div = m.total_div;
    }

    // same text layout as the reference's operator<< (object_model.h:55-63): it is the only place the
    // CLI emits the per-slice flow
    friend std::ostream &operator<<(std::ostream &os, const ObjectModel &M) {
        os << "C: (" << M.cx << ", " << M.cy << "); " << std::endl;
        os << "\t Shift: (" << M.dx << ", " << M.dy << "); " << " total: (" << M.total_dx << ", " << M.total_dy << ");" << std::endl;
        os << "\t Rot: " << M.rot << " total: " << M.total_rot << std::endl;
        os << "\t Div: " << M.div << " total: " << M.total_div << std::endl;
        os << "\t cnt: " << M.cnt << std::endl;
        return os;
    }
};

#endif  // BF_OBJECT_MODEL_H
