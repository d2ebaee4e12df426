// render.cu -- K6/K7 (placeholder until the renderer lands)
#include "common.cuh"
using namespace pnb;
extern "C" {
int pnb_render_image(int, int, const pnb_render_arrays *, double, double, double, double, double, double, double,
                     double, double, double, double, const float *, int, int, double, const double *, int,
                     const double *, int, float *, int)
{
    set_last_error("render_image: not built yet");
    return PNB_ERR_RUNTIME;
}
int pnb_render_grid(int, int, int, const pnb_render_arrays *, double, double, double, double, double, double, double,
                    double, const float *, int, int, double, const double *, int, const double *, int, const double *,
                    int, float *, int)
{
    set_last_error("render_grid: not built yet");
    return PNB_ERR_RUNTIME;
}
}
