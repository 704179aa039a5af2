// fluidsim_b200 -- command-line driver with the interface of the reference's src/main.cpp:
//
//   fluidsim_b200 <n_timesteps> <delta_t> <viscosity> <input_image> <velocity_field> [output_dir]
//
// Same argument checks and exit codes (main.cpp:105-140), same stdout lines (:214-215, :67, :250),
// same frames <output_dir>/<i>.png, timing mode when no output_dir is given (:110-113).  It drives
// the fluid.hpp entry points of the CUDA build (simulate_fluid_step / advect_color_step on device
// buffers, main.cpp:222,225), which fluid_shim.cpp forwards to libpfs_b200.so.
// Frames are converted to bytes on the device (pfs_image_to_rgba8, the reference's (png_byte)(x*255.0)), only those
// bytes cross PCIe (on a second stream), and the PNGs are encoded by worker threads while the next steps run
// (frame_writer.hpp; PFS_FRAME_WRITERS=0 gives the serial copy-and-encode of the reference).  Two deliberate differences from the reference's CUDA driver, both needed to
// reproduce what its CPU build computes: the temporary velocity buffer IS uploaded (main.cpp:203-210 never initialises
// d_vtmp although channel 2 of it is the first pressure guess), and the clock stops after the
// device is idle (main.cpp:247 stops it before cudaDeviceSynchronize, :252-255).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#include <cuda_runtime_api.h>

#define USE_CUDA
#include "../../include/pfs_b200.h"
#include "fluid.hpp"
#include "png_io.hpp"
#include "frame_writer.hpp"

#define NUM_CHANNELS (4)
#define VP_RANGE (2.0)

static void usage(const char *prog)
{
    std::cerr << "Usage: " << prog
              << " <n_timesteps> <delta_t> <viscosity> <input_image> <velocity_field> [output_dir]" << std::endl;
}

static void *pinned_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;   // utils.hpp:69-76
    return p;
}

static bool cuda_ok(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return true;
    std::cerr << what << ": " << cudaGetErrorString(e) << std::endl;
    return false;
}

int main(int argc, char *argv[])
{
    int flags = 0;
    if (argc != 6 && argc != 7) {
        usage(argv[0]);
        return 1;
    }
    if (argc == 6) flags |= 1;   // timing mode, no file writing
    int n_timesteps = atoi(argv[1]);
    float delta_t = atof(argv[2]);
    float viscosity = atof(argv[3]);
    if (n_timesteps <= 0) {
        std::cerr << "Timesteps must be greater than 0." << std::endl;
        return 1;
    }
    if (delta_t <= 0) {
        std::cerr << "Delta T must be greater than 0." << std::endl;
        return 1;
    }

    pngio::png_image input_image, velocity_image;
    vp_field image, vp, itmp, vtmp;
    if (pngio::read_png_to_array(&input_image, argv[4], &image.data, pinned_alloc) != 0) {
        std::cerr << "Something went wrong reading the input image..." << std::endl;
        return 1;
    }
    if (pngio::read_png_to_array(&velocity_image, argv[5], &vp.data, pinned_alloc) != 0) {
        std::cerr << "Something went wrong reading the initial velocity field..." << std::endl;
        return 1;
    }
    image.x = itmp.x = input_image.width;
    image.y = itmp.y = input_image.height;
    vp.x = vtmp.x = velocity_image.width;
    vp.y = vtmp.y = velocity_image.height;
    image.z = itmp.z = vp.z = vtmp.z = NUM_CHANNELS;

    // main.cpp:170-179: every channel v <- v*2.0 - 1.0 (double arithmetic, stored as float)
    const size_t vel_floats = (size_t)vp.x * vp.y * NUM_CHANNELS;
    const size_t img_floats = (size_t)image.x * image.y * NUM_CHANNELS;
    for (size_t i = 0; i < vel_floats; i++) {
        float v = vp.data[i];
        vp.data[i] = (v * VP_RANGE) - VP_RANGE / 2.0;
    }
    const size_t input_bytes = sizeof(float) * img_floats, velocity_bytes = sizeof(float) * vel_floats;

    // main.cpp:186-195: vtmp = (-1,-1,-1,+1) per cell
    vtmp.data = (float *)pinned_alloc(velocity_bytes);
    if (!vtmp.data) return 1;
    for (size_t i = 0; i < vel_floats; i++) vtmp.data[i] = ((i % 4) == 3) ? 1.0f : -1.0f;

    // frames leave the device as bytes: (png_byte)(x*255.0) of utils.hpp:129-131 is applied on the device, by the kernel
    // that advects the image (pfs_ctx_advect_color_step_rgba8) or, on the stateless path, by pfs_image_to_rgba8
    const size_t frame_bytes = img_floats;
    unsigned char *d_frame = nullptr, *h_frame = nullptr;
    int n_writers = (int)(std::thread::hardware_concurrency() / 2);   // PNG encoding is the slow part of a frame
    if (n_writers < 1) n_writers = 1;
    if (n_writers > 8) n_writers = 8;
    if (const char *e = getenv("PFS_FRAME_WRITERS")) n_writers = atoi(e);
    if (n_writers > 64) n_writers = 64;
    FrameWriter *writer = nullptr;
    if (flags == 0 && n_writers > 0) {
        writer = new FrameWriter(input_image, image.x, image.y, image.z, n_writers + 2, n_writers);
        if (!writer->ok()) {
            std::cerr << "cannot set up the frame writer" << std::endl;
            return 1;
        }
    } else if (flags == 0) {
        if (!cuda_ok(cudaMalloc((void **)&d_frame, frame_bytes), "cudaMalloc frame") ||
            !cuda_ok(cudaMallocHost((void **)&h_frame, frame_bytes), "cudaMallocHost frame"))
            return 1;
    }
    float *d_image = nullptr, *d_vp = nullptr, *d_itmp = nullptr, *d_vtmp = nullptr;
    if (!cuda_ok(cudaMalloc((void **)&d_image, input_bytes), "cudaMalloc") ||
        !cuda_ok(cudaMalloc((void **)&d_vp, velocity_bytes), "cudaMalloc") ||
        !cuda_ok(cudaMalloc((void **)&d_itmp, input_bytes), "cudaMalloc") ||
        !cuda_ok(cudaMalloc((void **)&d_vtmp, velocity_bytes), "cudaMalloc"))
        return 1;
    if (!cuda_ok(cudaMemcpy(d_image, image.data, input_bytes, cudaMemcpyHostToDevice), "H2D image") ||
        !cuda_ok(cudaMemcpy(d_vp, vp.data, velocity_bytes, cudaMemcpyHostToDevice), "H2D vp") ||
        !cuda_ok(cudaMemcpy(d_vtmp, vtmp.data, velocity_bytes, cudaMemcpyHostToDevice), "H2D vtmp"))
        return 1;

    std::cout << "Simulating [" << velocity_image.height << " x " << velocity_image.width << "] domain for "
              << n_timesteps << " timesteps at dt=" << delta_t << "..." << std::endl;

    // The state lives in a persistent context of the library (pfs_ctx_*: planar layout between steps, the interleaved
    // buffers above are only its upload source); PFS_DRIVER_STATELESS=1 drives the fluid.hpp entry points on the
    // caller-owned buffers instead, exactly as the reference's loop does (main.cpp:222,225).  Same results either way.
    const char *sl_env = getenv("PFS_DRIVER_STATELESS");
    const bool stateless = sl_env && sl_env[0] == '1';
    pfs_ctx *ctx = nullptr;
    if (!stateless) {
        if (pfs_ctx_create(&ctx, vp.x, vp.y, image.x, image.y) != PFS_OK ||
            pfs_ctx_upload(ctx, d_vp, d_vtmp, d_image, nullptr) != PFS_OK) {
            std::cerr << pfs_last_error() << std::endl;
            return 1;
        }
    }

    auto time_start = std::chrono::high_resolution_clock::now();
    for (int i = 0; i < n_timesteps; i++) {
        const float *frame_src = nullptr;
        std::string outpath;
        if (flags == 0) {
            outpath = std::string(argv[6]);
            if (!outpath.empty() && outpath.back() != '/') outpath += "/";
            outpath += std::to_string(i) + ".png";
        }
        bool frame_formed = false;
        if (ctx && flags == 0) {
            // frame mode: the kernel that advects the image also stores the frame's bytes (no separate pass over the image)
            unsigned char *dst = writer ? writer->begin(outpath) : d_frame;
            if (pfs_ctx_simulate_fluid_step(ctx, delta_t, viscosity, NUM_JACOBI_ITERS, NUM_JACOBI_ITERS, nullptr) != PFS_OK ||
                pfs_ctx_advect_color_step_rgba8(ctx, delta_t, dst, nullptr) != PFS_OK) {
                std::cerr << pfs_last_error() << std::endl;
                return 1;
            }
            frame_formed = true;
        } else if (ctx) {
            if (pfs_ctx_step(ctx, 1, delta_t, viscosity, NUM_JACOBI_ITERS, NUM_JACOBI_ITERS, nullptr) != PFS_OK) {
                std::cerr << pfs_last_error() << std::endl;
                return 1;
            }
        } else {
            simulate_fluid_step(&d_vp, &d_vtmp, delta_t, viscosity, vp.x, vp.y, vp.z);
            advect_color_step(&d_image, &d_itmp, &d_vp, delta_t, image.x, image.y, image.z, vp.x, vp.y, vp.z);
            frame_src = d_image;
        }
        if (flags == 0) {
            std::cout << "[" << i << "] Writing to : " << outpath << std::endl;
            if (writer) {
                if (!(frame_formed ? writer->commit() : writer->submit(frame_src, outpath))) {
                    std::cerr << writer->error() << std::endl;
                    return 1;
                }
                continue;
            }
            if (!frame_formed && pfs_image_to_rgba8(frame_src, d_frame, image.x, image.y, image.z, nullptr) != PFS_OK) {
                std::cerr << pfs_last_error() << std::endl;
                return 1;
            }
            if (!cuda_ok(cudaMemcpy(h_frame, d_frame, frame_bytes, cudaMemcpyDeviceToHost), "D2H frame")) return 1;
            if (pngio::write_png_from_bytes(&input_image, outpath.c_str(), h_frame) != 0)
                std::fprintf(stderr, "warning: cannot write %s\n", outpath.c_str());   // as the reference: not fatal (main.cpp:68)
        }
    }
    cudaDeviceSynchronize();
    if (writer && !writer->finish()) {   // every frame is on disk before the clock stops
        std::cerr << writer->error() << std::endl;
        return 1;
    }
    auto time_end = std::chrono::high_resolution_clock::now();
    std::cout << n_timesteps << " timesteps took "
              << std::chrono::duration_cast<std::chrono::microseconds>(time_end - time_start).count() << " us."
              << std::endl;

    delete writer;
    pfs_ctx_destroy(ctx);
    cudaFreeHost(image.data);
    cudaFreeHost(vp.data);
    cudaFreeHost(vtmp.data);
    cudaFree(d_image);
    cudaFree(d_vp);
    cudaFree(d_itmp);
    cudaFree(d_vtmp);
    if (d_frame) cudaFree(d_frame);
    if (h_frame) cudaFreeHost(h_frame);
    return 0;
}
