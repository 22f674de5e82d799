// ORACLE / TEST INFRASTRUCTURE.  Drives the reference's OWN intersection code (martingleich/RayTracingInAWeekend, src/lib) on rays
// from a file, so that oracle/ can be pinned against the real thing the day a Rust toolchain is available (tests/test_ref_harness.py).
//
//   rtiaw_ref_harness <in.bin> <out.bin>
//   in : u32 n_items | u32 kinds[n] (0 sphere, 1 rect, 2 triangle) | f32 geom[n][24] (layout of RtItem.geom, include/b200rt.h)
//        | u32 n_rays | f32 t_min | f32 t_max | f32 org[n_rays][3] | f32 dir[n_rays][3]
//   out: u32 hit_id[n_rays] (0xFFFFFFFF = miss) | f32 t[n_rays] | f32 normal[n_rays][3] | u8 front_face[n_rays]
//
// Every element gets its own Material, so the element a ray hit is recovered from HitInteraction::material (the reference's hit
// record carries no element index).  Scene = Scene::new(BoundingVolumeHierarchy::new(elements)), as NodeRef::finish builds it
// (src/app/worlds/world_builder.rs:282).
use std::fs;
use std::io::Write;

use rand::SeedableRng;
use ray_tracing_in_a_weekend::rect_geometry::{RectGeometry, RectPlane};
use ray_tracing_in_a_weekend::sphere_geometry::SphereGeometry;
use ray_tracing_in_a_weekend::triangle_geometry::TriangleGeometry;
use ray_tracing_in_a_weekend::{
    BoundingVolumeHierarchy, Dir3, Geometry, Material, Point3, Ray, Scene, SceneElement, Vec2f,
};

struct Reader<'a> {
    b: &'a [u8],
    p: usize,
}
impl<'a> Reader<'a> {
    fn u32(&mut self) -> u32 {
        let v = u32::from_le_bytes(self.b[self.p..self.p + 4].try_into().unwrap());
        self.p += 4;
        v
    }
    fn f32(&mut self) -> f32 {
        f32::from_bits(self.u32())
    }
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let data = fs::read(&args[1]).expect("cannot read input");
    let mut r = Reader { b: &data, p: 0 };
    let n = r.u32() as usize;
    let kinds: Vec<u32> = (0..n).map(|_| r.u32()).collect();
    let geom: Vec<[f32; 24]> = (0..n)
        .map(|_| {
            let mut g = [0.0f32; 24];
            for v in g.iter_mut() {
                *v = r.f32();
            }
            g
        })
        .collect();
    let n_rays = r.u32() as usize;
    let t_min = r.f32();
    let t_max = r.f32();
    let org: Vec<[f32; 3]> = (0..n_rays).map(|_| [r.f32(), r.f32(), r.f32()]).collect();
    let dir: Vec<[f32; 3]> = (0..n_rays).map(|_| [r.f32(), r.f32(), r.f32()]).collect();

    let materials: Vec<Material> = (0..n).map(|_| Material::Dielectric { index_of_refraction: 1.5 }).collect();
    let elements: Vec<SceneElement> = (0..n)
        .map(|i| {
            let g = &geom[i];
            let geo = match kinds[i] {
                0 => Geometry::Sphere(SphereGeometry::new(Point3::new(g[0], g[1], g[2]), g[3])),
                1 => Geometry::Rect(RectGeometry {
                    rect_plane: match g[0] as u32 {
                        0 => RectPlane::Xy,
                        1 => RectPlane::Xz,
                        _ => RectPlane::Yz,
                    },
                    dist: g[1],
                    r0: (g[2], g[3]),
                    r1: (g[4], g[5]),
                }),
                2 => Geometry::Triangle(TriangleGeometry {
                    positions: [Point3::new(g[0], g[1], g[2]), Point3::new(g[3], g[4], g[5]), Point3::new(g[6], g[7], g[8])],
                    normals: [Dir3::new(g[9], g[10], g[11]), Dir3::new(g[12], g[13], g[14]), Dir3::new(g[15], g[16], g[17])],
                    texture_coords: [Vec2f::new(g[18], g[19]), Vec2f::new(g[20], g[21]), Vec2f::new(g[22], g[23])],
                }),
                k => panic!("geometry kind {} is outside the harness", k),
            };
            SceneElement::SurfaceGeometry(geo, &materials[i])
        })
        .collect();
    let items: Vec<&SceneElement> = elements.iter().collect();
    let root = SceneElement::BoundingVolumeHierarchy(BoundingVolumeHierarchy::new(items, &(0.0..0.0)));
    let scene = Scene::new(&root);
    let mut rng = rand_xoshiro::Xoroshiro128PlusPlus::from_seed([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]);

    let base = materials.as_ptr() as usize;
    let stride = std::mem::size_of::<Material>();
    let mut hit_id = vec![0xFFFF_FFFFu32; n_rays];
    let mut t = vec![f32::INFINITY; n_rays];
    let mut nrm = vec![0.0f32; n_rays * 3];
    let mut ff = vec![0u8; n_rays];
    for i in 0..n_rays {
        // Ray::new debug-asserts a unit direction; build the struct directly so that adversarial rays do not abort a debug build
        let ray = Ray {
            origin: Point3::new(org[i][0], org[i][1], org[i][2]),
            direction: Dir3::new(dir[i][0], dir[i][1], dir[i][2]),
            time: 0.0,
        };
        if let Some(h) = scene.hit(&ray, &(t_min..t_max), &mut rng) {
            hit_id[i] = ((h.material as *const Material as usize - base) / stride) as u32;
            t[i] = h.t;
            nrm[3 * i] = h.normal.right();
            nrm[3 * i + 1] = h.normal.up();
            nrm[3 * i + 2] = h.normal.forward();
            ff[i] = h.front_face as u8;
        }
    }
    let mut out = fs::File::create(&args[2]).expect("cannot create output");
    for v in &hit_id {
        out.write_all(&v.to_le_bytes()).unwrap();
    }
    for v in &t {
        out.write_all(&v.to_bits().to_le_bytes()).unwrap();
    }
    for v in &nrm {
        out.write_all(&v.to_bits().to_le_bytes()).unwrap();
    }
    out.write_all(&ff).unwrap();
}
