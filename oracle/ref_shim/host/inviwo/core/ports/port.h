// stand-in (un-vendored Inviwo core): ppm/photondata.h includes it but uses nothing from it
#pragma once
