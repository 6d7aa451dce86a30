// Harness of tests/test_image_formats.py::test_decoders_survive_malformed_files: the ingest decoders (csrc/image_io.cpp,
// jpeg_io.cpp, image_formats.cpp) compiled with AddressSanitizer + UndefinedBehaviorSanitizer, run over mutated copies
// of the fixture corpus.  Every file named on the command line is decoded; decoded pixels are read once.
#include <cstdio>
#include <cstdlib>

#include "astc_b200.h"

extern "C" void astc_b200_free_host_buffer(void *p) { std::free(p); }      // lives in astc_capi.cu in the product library

int main(int argc, char **argv)
{
    int decoded = 0;
    for (int i = 1; i < argc; ++i) {
        int w = 0, h = 0, c = 0;
        uint8_t *img = nullptr;
        if (astc_b200_load_image(argv[i], i & 1, &w, &h, &c, &img) == 0) {
            volatile unsigned sum = 0;
            for (long k = 0; k < long(w) * h * 4; ++k) sum += img[k];
            astc_b200_free_host_buffer(img);
            ++decoded;
        }
    }
    std::printf("decoded %d of %d\n", decoded, argc - 1);
    return 0;
}
