// main.cu — the training driver the reference's README promises ("training loop", README.md:55) but
// whose src/main.cu is a 19-line tensor demo (src/main.cu:9-19).  Written against the reference's own
// tensor<T> API plus network<T> (include/network.cuh): the loop, the loss log, checkpoint / resume and
// the accuracy check live here; every step runs on the GPU inside libpinn_b200.so.
//
//   pinn_train [--config 1..5] [--points N] [--width H --depth L] [--precision fp32|fp16x3|bf16x6|bf16x3|bf16] [--steps N]
//              [--log-every K] [--resample-every K] [--csv file] [--save file] [--resume file] [--eval nodes_per_axis]
//
// Built into bin/pinn_train by the repo's Makefile next to the drop-in test (it needs the reference's
// tensor headers, which exist only where the reference checkout is mounted; the binary travels).
#include<chrono>
#include<cstdlib>
#include<fstream>
#include<iomanip>
#include<iostream>
#include<map>
#include<network.cuh>
#include<utils.cuh>

using std::cout;
using std::endl;
using network_f = network<float>;

namespace {
    struct options {
        int config = 2;
        size_t points = 0, steps = 1000, log_every = 100, resample_every = 0, eval = 0;
        int width = 0, depth = 0;
        std::string csv, save, resume, precision;
    };

    [[noreturn]] void usage(const std::string &why) {
        std::cerr << "pinn_train: " << why << "\n"
                  << "usage: pinn_train [--config 1..5] [--points N] [--width H --depth L] [--precision fp32|fp16x3|bf16x6|bf16x3|bf16]\n"
                  << "                  [--steps N] [--log-every K] [--resample-every K] [--csv file] [--save file] [--resume file]\n"
                  << "                  [--eval nodes_per_axis]" << endl;
        std::exit(2);
    }

    options parse(int argc, char **argv) {
        options o;
        std::map<std::string, std::string> kv;
        for(int i = 1; i < argc; i += 2) {
            const std::string key = argv[i];
            if (key.rfind("--", 0) != 0 || i + 1 >= argc) usage("expected --key value pairs, got '" + key + "'");
            kv[key.substr(2)] = argv[i + 1];
        }
        auto number = [&](const char *key, size_t fallback) -> size_t {
            const auto it = kv.find(key);
            if (it == kv.end()) return fallback;
            char *end = nullptr;
            const unsigned long long v = std::strtoull(it->second.c_str(), &end, 10);
            if (!end || *end) usage(std::string{"--"} + key + " needs a non-negative integer");
            kv.erase(it);
            return static_cast<size_t>(v);
        };
        auto text = [&](const char *key) -> std::string {
            const auto it = kv.find(key);
            if (it == kv.end()) return {};
            std::string v = it->second; kv.erase(it);
            return v;
        };
        o.config = static_cast<int>(number("config", 2));
        o.points = number("points", 0);
        o.width = static_cast<int>(number("width", 0));
        o.depth = static_cast<int>(number("depth", 0));
        o.steps = number("steps", 1000);
        o.log_every = number("log-every", 100);
        o.resample_every = number("resample-every", 0);
        o.eval = number("eval", 0);
        o.csv = text("csv"); o.save = text("save"); o.resume = text("resume"); o.precision = text("precision");
        if (!kv.empty()) usage("unknown option --" + kv.begin()->first);
        if (o.log_every == 0) usage("--log-every must be positive");
        return o;
    }
}

int main(int argc, char **argv) {
    const options opt = parse(argc, argv);
    try {
        pinn_config cfg = network_f::baseline(opt.config);
        if (opt.points) {
            cfg.n_interior = opt.points;
            cfg.n_boundary = opt.points / 32 > 256? opt.points / 32: 256;
            cfg.n_initial = cfg.pde == PINN_PDE_HELMHOLTZ? 0: cfg.n_boundary;
        }
        if (opt.width) cfg.width = opt.width;
        if (opt.depth) cfg.depth = opt.depth;
        if (!opt.precision.empty()) {
            static const std::map<std::string, int> known = {{"fp32", PINN_PREC_FP32}, {"fp16x3", PINN_PREC_FP16X3}, {"bf16x6", PINN_PREC_BF16X6},
                                                             {"bf16x3", PINN_PREC_BF16X3}, {"bf16", PINN_PREC_BF16}};
            const auto it = known.find(opt.precision);
            if (it == known.end()) usage("--precision must be one of fp32, fp16x3, bf16x6, bf16x3, bf16");
            cfg.precision = it->second;
        }

        network_f net(cfg);
        net.sample();
        size_t resumed_at = 0;                       // Adam step count of the checkpoint: the resample schedule continues from it
        if (!opt.resume.empty()) {
            net.load(opt.resume);
            resumed_at = net.loss().step;
            if (opt.resample_every && resumed_at >= opt.resample_every) {
                const size_t epoch = resumed_at / opt.resample_every;      // the point sets the uninterrupted run would be on
                net.resample(cfg.seed_interior + 16 * epoch, cfg.seed_boundary + 16 * epoch, cfg.seed_initial + 16 * epoch);
            }
        }

        cout << "pinn_train: pde " << cfg.pde << ", " << cfg.d_in << " -> [" << cfg.width << " x " << cfg.depth << "] -> " << cfg.d_out
             << ", " << net.size() << " parameters, " << net.channels() << " Taylor channels, "
             << net.count(point_set::interior) << " interior / " << net.count(point_set::boundary) << " boundary / "
             << net.count(point_set::initial) << " initial points" << endl;

        std::ofstream csv;
        if (!opt.csv.empty()) {
            csv.open(opt.csv);
            if (!csv) throw std::runtime_error{"main: cannot open " + opt.csv};
            csv << "step,loss,mse_residual,mse_boundary,mse_initial,seconds\n";
        }

        const auto t0 = std::chrono::steady_clock::now();
        size_t done = 0;
        step_result last{};
        while(done < opt.steps) {
            // steps between two log lines run back to back on the device (graph replay), one loss read per chunk
            size_t chunk = std::min(opt.log_every, opt.steps - done);
            const size_t at = resumed_at + done;     // global step count: a resumed run resamples where the uninterrupted one would
            if (opt.resample_every) chunk = std::min(chunk, opt.resample_every - at % opt.resample_every);
            last = net.train(chunk);
            done += chunk;
            const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (done % opt.log_every == 0 || done == opt.steps) {
                cout << "step " << std::setw(8) << last.step << "  loss " << std::scientific << std::setprecision(6) << last.loss
                     << "  residual " << last.mse_residual << "  boundary " << last.mse_boundary << "  initial " << last.mse_initial
                     << std::defaultfloat << "  " << std::setprecision(4) << seconds << " s" << endl;
                if (csv.is_open())
                    csv << last.step << ',' << std::setprecision(9) << last.loss << ',' << last.mse_residual << ',' << last.mse_boundary << ','
                        << last.mse_initial << ',' << seconds << '\n';
            }
            if (opt.resample_every && (resumed_at + done) % opt.resample_every == 0) {
                const size_t epoch = (resumed_at + done) / opt.resample_every;        // fresh collocation sets: seeds advance by epoch
                net.resample(cfg.seed_interior + 16 * epoch, cfg.seed_boundary + 16 * epoch, cfg.seed_initial + 16 * epoch);
            }
        }
        const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (opt.steps)
            cout << "trained " << opt.steps << " steps in " << std::setprecision(4) << seconds << " s: "
                 << std::setprecision(6) << static_cast<double>(opt.steps) * static_cast<double>(cfg.n_interior) / seconds << " points/s" << endl;

        if (!opt.save.empty()) { net.save(opt.save); cout << "saved " << opt.save << endl; }

        if (opt.eval) {
            const pinn_eval_out e = net.evaluate(opt.eval);
            cout << "grid " << e.n << " nodes: residual mse " << std::scientific << std::setprecision(6) << e.mse_residual;
            for(int o = 0; o < cfg.d_out; o++)
                if (e.rel_l2[o] == e.rel_l2[o]) cout << "  rel_l2[" << o << "] " << e.rel_l2[o] << "  max_abs[" << o << "] " << e.max_abs[o];
            cout << std::defaultfloat << endl;
        }
    }
    catch(const std::exception &e) {
        std::cerr << "pinn_train: " << e.what() << endl;
        return 1;
    }
    return 0;
}
