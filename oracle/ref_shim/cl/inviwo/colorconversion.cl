// colorconversion.cl (Inviwo, un-vendored) -- stand-in: sRGB (D65) -> XYZ -> CIE L*a*b*; arithmetic = oracle/orc_grid.c rgb2lab
#ifndef COLORCONVERSION_CL
#define COLORCONVERSION_CL
CLC_INLINE float3 rgb2lab(float3 rgb) {
    float c[3] = {rgb.x, rgb.y, rgb.z}, lin[3];
    for (int k = 0; k < 3; ++k) lin[k] = c[k] > 0.04045f ? cpm_powf((c[k] + 0.055f) / 1.055f, 2.4f) : c[k] / 12.92f;
    float X = 0.4124564f * lin[0] + 0.3575761f * lin[1] + 0.1804375f * lin[2];
    float Y = 0.2126729f * lin[0] + 0.7151522f * lin[1] + 0.0721750f * lin[2];
    float Z = 0.0193339f * lin[0] + 0.1191920f * lin[1] + 0.9503041f * lin[2];
    float xyz[3] = {X / 0.95047f, Y / 1.0f, Z / 1.08883f}, f[3];
    for (int k = 0; k < 3; ++k) f[k] = xyz[k] > 0.008856f ? cpm_cbrtf(xyz[k]) : 7.787f * xyz[k] + 16.0f / 116.0f;
    return make_float3(116.0f * f[1] - 16.0f, 500.0f * (f[0] - f[1]), 200.0f * (f[1] - f[2]));
}
#endif
