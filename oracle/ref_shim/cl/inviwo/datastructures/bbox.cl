// datastructures/bbox.cl (Inviwo, un-vendored) -- stand-in: uploaded as 2 x vec4 (ppm/processor/progressivephotontracercl.cpp:192-195)
#ifndef BBOX_CL
#define BBOX_CL
typedef struct BBox {
    float3 pMin;
    float3 pMax;
} BBox;
#endif
