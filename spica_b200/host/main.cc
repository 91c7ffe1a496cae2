// `spica -i scene.xml [-t N] [-o OUT]`: the reference's render entry point (spica/main.cc:20-55)
// on the GPU host. -t is accepted and ignored (there are no CPU render threads); --gpus N (or
// SPICA_GPUS) partitions the samples over N GPUs; --seed fixes the sampler key; --save-passes
// restores the reference's save-after-every-pass behaviour.
#include <cstring>
#include <filesystem>
#include <iostream>

#include "core.h"

namespace fs = std::filesystem;
using namespace spica;

int main(int argc, char** argv) {
    hostPhaseLap("main entered");
    std::string input, output;
    int threads = 4;
    HostOptions& opt = hostOptions();
    if (const char* g = getenv("SPICA_GPUS")) opt.gpus = atoi(g);
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) { std::cout << "missing value for " << a << std::endl; exit(1); } return argv[++i]; };
        if (a == "-i" || a == "--input") input = next();
        else if (a == "-t" || a == "--threads") threads = atoi(next().c_str());
        else if (a == "-o" || a == "--output") output = next();
        else if (a == "--gpus") opt.gpus = atoi(next().c_str());
        else if (a == "--seed") opt.seed = strtoull(next().c_str(), nullptr, 10);
        else if (a == "--save-passes") opt.savePasses = true;
        else { std::cout << "usage: spica -i scene.xml [-t N] [-o OUT] [--gpus N] [--seed S] [--save-passes]" << std::endl; return 1; }
    }
    if (input.empty()) { std::cout << "usage: spica -i scene.xml [-t N] [-o OUT] [--gpus N] [--seed S] [--save-passes]" << std::endl; return 1; }
    if (output.empty()) {                                               // main.cc:36-42
        std::error_code ec;
        const std::string full = fs::canonical(fs::absolute(fs::path(input)), ec).string();
        if (ec) FatalError("Failed to open file:%s\n", input.c_str());
        output = full.substr(0, full.find_last_of('.'));
    }
    hostGpuPrewarm(opt.gpus);                                           // CUDA start-up and the communicator run behind the scene loading
    RenderParams& params = RenderParams::getInstance();
    params.add("numUserThreads", threads);
    params.add("outputFile", output);
    SceneParser parser(input);
    parser.parse();
    hostGpuFinish();
    hostPhaseLap("parse() returned");
    return 0;
}
