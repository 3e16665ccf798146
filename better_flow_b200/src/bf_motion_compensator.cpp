// bf_motion_compensator -- command-line front end, drop-in for the reference tool
// (reference: better_flow_core/src/bf_motion_compensator.cpp).  Every flag of the reference is accepted
// with the same meaning and the same stdout formats; the compile-time constants of the reference
// (EVENT_WIDTH, TIME_WIDTH, RES_X/RES_Y, scale, max_iter) are additionally exposed as flags whose
// defaults reproduce the reference.  The computation always runs on the CUDA back-end.
#include <better_flow/common.h>
#include <better_flow/dvs_flow.h>
#include <better_flow/opencl_driver.h>

#include <chrono>

#define EVENT_WIDTH 50000
#define TIME_WIDTH 0.2

static float time_refresh = 0.033f;
static unsigned long long int event_refresh = 20000;

static bool manual = false;
static bool quiet = false;
static char *file = NULL;
static char *outFileName = NULL;
static bool gpu = false;
static bool img = false;
static bool video = false;
static bool stm_disable = false;
static bool bufferize_file = false;
static std::string img_prefix = "./";
static std::string video_name = "./out.avi";
static int video_fps = 60;
// additions
static int max_iter = -1;
static int scale = 3;
static double slice_time = TIME_WIDTH;
static long long max_events = EVENT_WIDTH;
static int sensor_w = 240, sensor_h = 180;
static char *flowOutName = NULL;
static int batch = 1;
static int gpus = 1;
static bool device_ring = true;
static bool optimizer_local = false;
static int device = 0;

static void lPrintVersion() {
    printf("DVS flow estimator (better flow), %s (build %s @ %s)\n", BF_VERSION, __DATE__, __TIME__);
    printf("\tDefault maximum event memory of %i events\n\tand slice size of %f seconds (both run-time options here).\n",
           EVENT_WIDTH, TIME_WIDTH);
}

static void usage(int ret) {
    lPrintVersion();
    printf("\nusage: bf_motion_compensator\n");
    printf("    [--refresh-time={0.0 - inf}]\t\tRun processing when at least this amount of time (floatimg point,\n");
    printf("                                \t\tseconds) has passed since the last processing, (default = %f)\n", time_refresh);
    printf("    [--refresh-event-count={0 - inf}]\t\tRun processing when at least this number of new events has\n");
    printf("                                     \t\tarrived since the last processing (default = %llu)\n", event_refresh);
    printf("    [-i/--interactive]\tEnable interactive mode (ignored: GUI feature)\n");
    printf("    [-G]\t\t\t\tUse GPU support (always on: the CUDA back-end is the only implementation)\n");
    printf("    [--stm-disable]\t\t\t\tDo not use previous estimate as a starting point for a new estimate\n");
    printf("    [--img]\t\t\t\tOutput flow images after every iteration (frame_N.ppm + frame_N.txt)\n");
    printf("    [--img-prefix <name>]\t\t\t\tSpecify prefix for the generated image files (default = %s)\n", img_prefix.c_str());
    printf("    [--video]\t\t\t\tOutput a video with flow frames (uncompressed YUV4MPEG2, 4:4:4)\n");
    printf("    [--video-name <name>]\t\t\t\tSpecify the name of the video file (default = %s)\n", video_name.c_str());
    printf("    [--video-fps=<value>]\t\t\t\tSpecify video framerate (default = %i)\n", video_fps);
    printf("    [--bufferize-file]\t\t\t\tRead input file to the buffer first (useful for performance testing)\n");
    printf("    [--quiet]\t\t\t\tSuppress the per-slice model dump\n");
    printf("    [-o <name>/--outfile=<name>]\tOutput filename (may be \"-\" for standard output)\n");
    printf("    [--version]\t\t\t\tPrint better flow version\n");
    printf("  additions (defaults reproduce the reference):\n");
    printf("    [--max-iter=N]\t\t\tMaximum number of optimisation steps, -1 = until convergence (default = %i)\n", max_iter);
    printf("    [--scale=N]\t\t\t\tImage scale 1/3/5 (default = %i)\n", scale);
    printf("    [--slice-time=SEC]\t\t\tTime span of the event buffer (default = %f)\n", slice_time);
    printf("    [--max-events=N]\t\t\tCapacity of the event buffer (default = %lld)\n", max_events);
    printf("    [--sensor=WxH]\t\t\tSensor size in pixels (default = %ix%i)\n", sensor_w, sensor_h);
    printf("    [--flow-out=<name>]\t\t\tWrite one line per slice: id n iters rc total_dx total_dy total_rot total_div cx cy dx dy rot div cnt\n");
    printf("    [--batch=N]\t\t\t\tWith --stm-disable: minimise N slices per kernel launch (default = %i)\n", batch);
    printf("    [--no-device-ring]\t\t\tKeep the slice ring on the host and hand every slice over in full (default without -o: the ring\n");
    printf("              \t\t\t\tlives on the device, only new events are uploaded, warm starts chain on the device)\n");
    printf("    [--gpus=N]\t\t\t\tWith --stm-disable and --batch: deal every batch to N devices (starting at --device),\n");
    printf("              \t\t\t\tone launch per device and one NCCL all-gather of the per-slice flow (default = %i)\n", gpus);
    printf("    [--optimizer=rolling|local]\tPer-slice optimiser: OptimizerRolling (default, what the reference tool runs) or\n");
    printf("                               \tOptimizerLocal (contrast-driven nx, ny descent; the slice model carries -nx, -ny)\n");
    printf("    [--device=N]\t\t\tCUDA device (default = %i)\n", device);
    printf("    <file to process or \"-\" for stdin>\n");
    exit(ret);
}

// (GCC treats `main` itself as cold code -- run once -- and inlines nothing into it; the per-event loop lives here.)
__attribute__((hot)) static int tool_main(int argc, char *argv[]) {
    if (argc == 1) usage(1);
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--help")) usage(0);
        else if (!strcmp(argv[i], "-v") || !strcmp(argv[i], "--version")) { lPrintVersion(); return 0; }
        else if (!strcmp(argv[i], "--quiet")) quiet = true;
        else if (!strncmp(argv[i], "--refresh-time=", 15)) time_refresh = atof(argv[i] + 15);
        else if (!strncmp(argv[i], "--refresh-event-count=", 22)) event_refresh = atoi(argv[i] + 22);
        else if (!strcmp(argv[i], "-G")) gpu = true;
        else if (!strcmp(argv[i], "-i") || !strcmp(argv[i], "--interactive")) manual = true;
        else if (!strcmp(argv[i], "--bufferize-file")) bufferize_file = true;
        else if (!strcmp(argv[i], "--stm-disable")) stm_disable = true;
        else if (!strcmp(argv[i], "--img")) img = true;
        else if (!strcmp(argv[i], "--img-prefix")) {
            if (++i == argc) { fprintf(stderr, "No output file specified after --img-prefix option.\n"); usage(1); }
            img_prefix = argv[i];
        }
        else if (!strcmp(argv[i], "--video")) video = true;
        else if (!strcmp(argv[i], "--video-name")) {
            if (++i == argc) { fprintf(stderr, "No output file specified after --video-name option.\n"); usage(1); }
            video_name = argv[i];
        }
        else if (!strncmp(argv[i], "--video-fps=", 12)) video_fps = atoi(argv[i] + 12);
        else if (!strcmp(argv[i], "-o")) {
            if (++i == argc) { fprintf(stderr, "No output file specified after -o option.\n"); usage(1); }
            outFileName = argv[i];
        }
        else if (!strncmp(argv[i], "--outfile=", 10)) outFileName = argv[i] + strlen("--outfile=");
        else if (!strncmp(argv[i], "--max-iter=", 11)) max_iter = atoi(argv[i] + 11);
        else if (!strncmp(argv[i], "--scale=", 8)) scale = atoi(argv[i] + 8);
        else if (!strncmp(argv[i], "--slice-time=", 13)) slice_time = atof(argv[i] + 13);
        else if (!strncmp(argv[i], "--max-events=", 13)) max_events = atoll(argv[i] + 13);
        else if (!strncmp(argv[i], "--sensor=", 9)) {
            if (sscanf(argv[i] + 9, "%dx%d", &sensor_w, &sensor_h) != 2) { fprintf(stderr, "Bad --sensor=WxH.\n"); usage(1); }
        }
        else if (!strncmp(argv[i], "--flow-out=", 11)) flowOutName = argv[i] + 11;
        else if (!strncmp(argv[i], "--batch=", 8)) batch = atoi(argv[i] + 8);
        else if (!strncmp(argv[i], "--gpus=", 7)) gpus = atoi(argv[i] + 7);
        else if (!strcmp(argv[i], "--no-device-ring")) device_ring = false;
        else if (!strcmp(argv[i], "--optimizer=local")) optimizer_local = true;
        else if (!strcmp(argv[i], "--optimizer=rolling")) optimizer_local = false;
        else if (!strncmp(argv[i], "--device=", 9)) device = atoi(argv[i] + 9);
        else if (!strcmp(argv[i], "-")) {}
        else if (argv[i][0] == '-') { fprintf(stderr, "Unknown option \"%s\".\n", argv[i]); usage(1); }
        else {
            if (file != NULL) {
                fprintf(stderr, "Multiple input files specified on command line: \"%s\" and \"%s\".\n", file, argv[i]);
                usage(1);
            }
            else file = argv[i];
        }
    }
    if (file == NULL) { fprintf(stderr, "No input file.\n"); usage(1); }
    if (scale != 1 && scale != 3 && scale != 5) { fprintf(stderr, "--scale must be 1, 3 or 5.\n"); return 1; }
    if (batch > 1 && !stm_disable) { fprintf(stderr, "--batch needs --stm-disable (warm-started slices form a chain).\n"); return 1; }
    if (gpus > 1 && !stm_disable) { fprintf(stderr, "--gpus needs --stm-disable (warm-started slices form a chain).\n"); return 1; }
    if (gpus > 1 && batch < gpus) { fprintf(stderr, "--gpus=%d needs --batch of at least %d slices.\n", gpus, gpus); return 1; }
    (void)gpu; (void)img; (void)video;

    const bool timing = std::getenv("BF_TIMING") != nullptr;
    const auto wall_start = std::chrono::steady_clock::now();
    auto stamp = [&](const char *what) {
        if (timing)
            std::cerr << "[timing] " << what << " at " << std::chrono::duration<double>(std::chrono::steady_clock::now() - wall_start).count()
                      << " s" << std::endl;
    };
    bf::set_sensor(sensor_h, sensor_w);   // RES_X = rows, RES_Y = columns
    OpenCLDriver::init(device);           // same call site as the reference's -G branch (:132-133)
    stamp("device selected");

    DVS_flow<EVENT_WIDTH, FROM_SEC(TIME_WIDTH)> estimator(event_refresh, FROM_SEC(time_refresh), 0, (size_t)max_events,
                                                          (sll)FROM_SEC(slice_time));
    if (outFileName != NULL) estimator.set_accumulate();
    else if (img || video) {
        // frames are made from the per-event warped positions: keep reading them back
    } else {
        estimator.set_lazy_events(true);    // nothing reads the per-event flow: skip its read-back (the per-slice models are unaffected)
        estimator.set_device_ring(device_ring);
    }
    if (manual) estimator.set_manual_mode(true);
    if (img) estimator.set_generate_pictures(true, img_prefix);
    if (video) estimator.set_generate_video(true, video_name, video_fps);
    if (stm_disable) estimator.set_stm_disable(true);
    estimator.set_max_iter(max_iter);
    estimator.set_scale(scale);
    estimator.set_batch(batch);
    estimator.set_gpus(gpus);
    if (optimizer_local) {
        estimator.set_optimizer_local(true);
    }
    estimator.set_quiet(quiet);
    std::ofstream flow_out;
    if (flowOutName != NULL) {
        flow_out.open(flowOutName);
        estimator.set_flow_out(&flow_out);
    }

    estimator.prepare();                  // device context (and slice ring) now, not inside the first slice
    stamp("device objects created");
    const auto wall0 = std::chrono::steady_clock::now();
    bool final_done = false;
    const size_t flen = strlen(file);
    const bool binary = flen > 4 && !strcmp(file + flen - 4, ".bin");
    if (binary) bufferize_file = true;   // binary input is always read up front
    if (bufferize_file) {
        LinearEventCloud ec;
        std::vector<EventFile::BinaryRecord> recs;   // binary input stays in its 16-byte records; Events are built on the fly
        if (binary) recs = EventFile::read_binary_records(file);
        else EventFile::from_file(&ec, file);
        const ull n_total = binary ? (ull)recs.size() : (ull)ec.size();

        clock_t begin = std::clock();
        clock_t begin_slice = std::clock();
        const auto proc0 = std::chrono::steady_clock::now();
        ull i = 0;
        auto feed = [&](Event &e) {
            ++i;
            bool processed = estimator.add_event(e);
            if (processed) {
                clock_t end_slice = std::clock();
                std::cout << float(i * 100) / float(n_total) << " %\t" << i << "\t"
                          << (double(end_slice - begin_slice) / CLOCKS_PER_SEC) << " sec\t" << estimator.get_buf_size()
                          << " events\t" << double(estimator.get_time_diff()) / 1000000000.0 << " slice_td\t"
                          << double(estimator.get_buf_time_diff()) / 1000000000.0 << " buffer_td\n";
                begin_slice = std::clock();
            }
        };
        if (binary) {
            for (const auto &r : recs) {
                Event e(r.y, r.x, ull(r.t_ns));
                feed(e);
            }
        } else {
            for (auto &e : ec) feed(e);
        }
        clock_t end = std::clock();
        std::cout << "Toatal flow elapsed: " << double(end - begin) / CLOCKS_PER_SEC << " sec." << std::endl << std::flush;
        const double host_loop_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - proc0).count();
        if (timing) {
            std::cerr << "[timing] add_event loop returned after " << host_loop_s << " s (with the device ring: slices are only enqueued)" << std::endl;
            // wall time of the processing proper (events already in memory): add_event loop + the final slice + every
            // model read back.  (The reference's figure above is std::clock(): CPU time, summed over threads.)
            estimator.recompute();
            estimator.flush();
            final_done = true;
            const double w = std::chrono::duration<double>(std::chrono::steady_clock::now() - proc0).count();
            double tp = 0, ts = 0, tr = 0;
            estimator.ring_host_seconds(tp, ts, tr);
            if (tp + ts + tr > 0)
                std::cerr << "[timing] device ring, host seconds inside bf_ring_push " << tp << ", bf_ring_slice " << ts
                          << ", bf_ring_result (waiting for the GPU) " << tr << "; ring_slice() as a whole " << estimator.recompute_host_seconds() << std::endl;
            std::cerr << "[timing] processing " << i << " events in " << w << " s = " << double(i) / w / 1e6 << " Mev/s, slices "
                      << estimator.slices_done() << std::endl;
        }
    } else {
        std::cout << "Reading from file... (" << file << ")" << std::endl << std::flush;
        // block reader + from_chars (same values as `ifstream >>`), parsing one block ahead on a second thread
        PrefetchingTextEventReader event_file(file);
        ull i = 0;
        double t = 0;
        uint x = 0, y = 0;
        bool p = false;
        double t_0 = 0;   // the earliest timestamp in the file
        if (event_file.next(t_0, x, y, p)) {
            ++i;
            Event e(y, x, FROM_SEC(0));
            estimator.add_event(e);
        }
        while (event_file.next(t, x, y, p)) {
            t -= t_0;
            ++i;
            Event e(y, x, FROM_SEC(t));
            estimator.add_event(e);
        }
        std::cout << "Read and processed " << i << " events" << std::endl << std::flush;
    }

    stamp("stream consumed");
    if (!final_done) estimator.recompute();   // ensure that *every* event has been processed
    estimator.flush();
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
    if (!quiet)
        std::cerr << "slices " << estimator.slices_done() << ", slice-events " << estimator.events_done() << ", GD steps "
                  << estimator.iterations_done() << ", wall " << wall << " s" << std::endl;

    if (outFileName != NULL) {
        LinearEventCloudTemplate<Event> accumulated = estimator.get_accumulated();
        EventFile::to_file_uv(&accumulated, outFileName);
    }
    stamp("outputs written");
    CudaDriver::shutdown();
    stamp("shutdown");
    return 0;
}

int main(int argc, char *argv[]) { return tool_main(argc, argv); }
