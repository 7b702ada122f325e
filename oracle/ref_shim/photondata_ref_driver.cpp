// photondata_ref_driver.cpp -- C entry points over the reference's own ppm/photondata.cpp (compiled where it lies):
// progressive radius / irradiance scale of PhotonData and the host twin of the direction codec.  TEST INFRASTRUCTURE.
#include <modules/progressivephotonmapping/photondata.h>
#define REF_API extern "C" __attribute__((visibility("default")))
using namespace inviwo;

// out = {radius after `iterations` advanceToNextIteration(alpha) calls, getRadiusRelativeToSceneSize, getRelativeIrradianceScale,
//        iteration}
REF_API void ref_photondata_progress(size_t n_photons, int max_interactions, double radius_rel, double scene_radius, int iterations,
                                     double alpha, double out[4]) {
    PhotonData d;
    d.setSize(n_photons, max_interactions);
    d.setRadius(radius_rel, scene_radius);
    for (int i = 0; i < iterations; ++i) d.advanceToNextIteration(alpha);
    out[0] = d.getRadius();
    out[1] = d.getRadiusRelativeToSceneSize();
    out[2] = d.getRelativeIrradianceScale();
    out[3] = d.iteration();
}
REF_API double ref_photondata_sphere_volume(double r) { return PhotonData::sphereVolume(r); }
REF_API double ref_photondata_progressive_radius(double r, int it, double alpha) { return PhotonData::progressiveSphereRadius(r, it, alpha); }
REF_API void ref_photondata_constants(double out[4]) {
    out[0] = PhotonData::defaultRadiusRelativeToSceneRadius;
    out[1] = PhotonData::defaultSceneRadius;
    out[2] = PhotonData::scaleToMakeLightPowerOfOneVisibleForDirectionalLightSource;
    out[3] = PhotonData::defaultNumberOfPhotons;
}
REF_API size_t ref_photondata_number_of_photons(size_t n_photons, int max_interactions) {
    PhotonData d;
    d.setSize(n_photons, max_interactions);
    return d.getNumberOfPhotons();
}
REF_API void ref_photon_encode_direction(const float dir[3], float out[2]) {
    Photon p;
    p.setDirection(vec3(dir[0], dir[1], dir[2]));
    out[0] = p.encodedDirection.x;
    out[1] = p.encodedDirection.y;
}
REF_API void ref_photon_decode_direction(const float enc[2], float out[3]) {
    Photon p;
    p.encodedDirection = vec2(enc[0], enc[1]);
    vec3 d = p.getDirection();
    out[0] = d.x; out[1] = d.y; out[2] = d.z;
}
