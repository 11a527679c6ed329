// Test-infrastructure shim: forward to the nlohmann/json copy vendored in the image's cudnn_frontend
// (3.11.3; the reference pins 3.12.0 -- the API it uses is identical).  Path supplied by -I in the Makefile.
#pragma once
#include <cudnn_frontend/thirdparty/nlohmann/json.hpp>
