// TEST INFRASTRUCTURE ONLY (oracle/_ref): exposes the reference's vendored GLM 0.9.9.1 gtc/noise (dependencies/glm) through a C ABI,
// compiled with the same flags as the reference makefile (-O3, no -march => no FMA contraction).
#include <glm/gtc/noise.hpp>
extern "C" {
float ref_glm_simplex2(float x, float y) {return glm::simplex(glm::vec2(x, y));}
float ref_glm_perlin2 (float x, float y) {return glm::perlin (glm::vec2(x, y));}
float ref_glm_simplex3(float x, float y, float z) {return glm::simplex(glm::vec3(x, y, z));}
float ref_glm_perlin3 (float x, float y, float z) {return glm::perlin (glm::vec3(x, y, z));}
}
