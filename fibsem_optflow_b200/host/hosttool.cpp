// hosttool -- exposes the job driver's host-side pieces (JSON reader/writer, image readers, float
// TIFF writer) on the command line so that the CPU test-suite can check them without a GPU.
//   hosttool json <file>            parse (comments / trailing commas tolerated), print canonical dump
//   hosttool image <file>           decode PNG / PGM / TIFF to 8-bit grey; print "w h" then raw bytes
//   hosttool tiff <w> <h> <out>     read w*h floats from stdin, write a float TIFF
#include <cstdio>
#include <fstream>
#include <iostream>
#include <iterator>
#include <sstream>

#include "imageio.h"
#include "minijson.h"

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: hosttool json|image|tiff ...\n"); return 2; }
    const std::string cmd = argv[1];
    if (cmd == "json") {
        std::ifstream f(argv[2], std::ios::binary);
        std::stringstream ss;
        ss << f.rdbuf();
        try {
            std::cout << mj::dump(mj::parse(ss.str())) << "\n";
        } catch (const std::exception& e) {
            std::fprintf(stderr, "parse error: %s\n", e.what());
            return 1;
        }
        return 0;
    }
    if (cmd == "image") {
        imio::Gray8 img;
        std::string err;
        if (!imio::read_gray8(argv[2], img, err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }
        std::printf("%d %d\n", img.w, img.h);
        std::fwrite(img.px.data(), 1, img.px.size(), stdout);
        return 0;
    }
    if (cmd == "tiff" && argc == 5) {
        const int w = std::atoi(argv[2]), h = std::atoi(argv[3]);
        std::vector<float> v((size_t)w * h);
        if (std::fread(v.data(), 4, v.size(), stdin) != v.size()) return 1;
        return imio::write_tiff_f32(argv[4], v.data(), w, h) ? 0 : 1;
    }
    return 2;
}
