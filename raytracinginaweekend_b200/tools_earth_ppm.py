"""python -m raytracinginaweekend_b200.tools_earth_ppm [out.ppm]

Writes assets/earthmap.ppm (binary P6) from assets/earthmap.jpg for the C++ CLI (rtiaw_app), which reads image textures as PPM:
JPEG decoding is caller-side here, as it is in the reference (the `image` crate, src/app/worlds/world_builder.rs:19-31)."""
import os
import sys

from .api import ASSETS, earth_rgb8


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ASSETS, "earthmap.ppm")
    img = earth_rgb8()
    with open(out, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(img.tobytes())
    print(out)


if __name__ == "__main__":
    main()
