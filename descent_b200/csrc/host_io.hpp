// Host-side data front ends of the reference's examples; see host_io.cpp.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace descent {

std::vector<uint8_t> read_file_bytes(const std::string& path);
std::vector<uint8_t> gunzip(const uint8_t* data, size_t size);
std::vector<uint8_t> load_gz_bytes(const std::string& path);  // examples/fashion_mnist/main.rs:13-19

struct IdxImagesInfo { uint32_t images, rows, cols; size_t data_offset; };
struct IdxLabelsInfo { uint32_t items; size_t data_offset; };
IdxImagesInfo read_images_info(const uint8_t* bytes, size_t size);  // main.rs:26-33
IdxLabelsInfo read_labels_info(const uint8_t* bytes, size_t size);  // main.rs:35-40
// out[i, :] = image indices[i] as byte / 255 (main.rs:42-60); out[i] = label indices[i] as f32 (main.rs:62-72)
void unpack_images(const uint8_t* bytes, size_t size, const uint64_t* indices, size_t count, float* out);
void unpack_labels(const uint8_t* bytes, size_t size, const uint64_t* indices, size_t count, float* out);

struct JpegImage {
    int width = 0, height = 0;
    std::vector<uint8_t> rgb;  // [height, width, 3]
};
JpegImage decode_jpeg_rgb(const uint8_t* data, size_t size);  // examples/image_fit/main.rs:278-282 (stbi_load, Channels::Rgb)
void write_ppm(const std::string& path, const float* rgb, int width, int height);  // main.rs:421-435, as PPM

}  // namespace descent
