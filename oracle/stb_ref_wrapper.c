/* Checker for astc_encoder_b200/csrc/image_io.cpp + jpeg_io.cpp + image_formats.cpp: the reference's vendored
 * stb_image v2.22, compiled WHERE IT LIES in the reference checkout (never copied into this repo) into
 * oracle/_ref/libstb_ref.so by `make -C oracle stb_ref`.  TEST INFRASTRUCTURE ONLY, and only where the
 * reference checkout exists (the build container); elsewhere the committed fixtures under tests/golden/
 * made with it (tools/make_image_fixtures.py) stand in. */
#define STB_IMAGE_IMPLEMENTATION
#include STB_IMAGE_PATH
