// The stb implementation units the reference instantiates in its UI translation unit (Main.cpp:45-48), which the headless build
// does not compile; tinygltf needs the image-write symbols as well.
#define STB_IMAGE_IMPLEMENTATION
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include "stb_image.h"
#include "stb_image_write.h"
