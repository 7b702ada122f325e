// datastructures/lightsource.cl (Inviwo modules/opencl, un-vendored) -- stand-in: field use visible at isc/cl/light/light.cl:84-121
#ifndef LIGHTSOURCE_CL
#define LIGHTSOURCE_CL
typedef struct LightSource {
    float16 tm;
    float3 radiance;
    int type;
    float2 size;
    float area;
    float cosFOV;
} LightSource;
// inviwo::LightSourceType
#define LIGHT_AREA 0
#define LIGHT_CONE 1
#define LIGHT_POINT 2
#define LIGHT_DIRECTIONAL 3
#endif
