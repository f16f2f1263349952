// Instantiates the stb implementations the reference normally defines in Main.cpp:45-48
// (Main.cpp itself is UI code and is not built).  TEST INFRASTRUCTURE.
#define STB_IMAGE_IMPLEMENTATION
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include "stb_image.h"
#include "stb_image_write.h"
