--  gcore_b200.ads -- thin Ada binding of libgcore_b200.so (include/gcore.h).
--
--  NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no GNAT.  It is the binding a maintainer of
--  FROL256/ada-ray-tracer adds next to scene_hydra_embree.adb; every Import names the C entry point it binds and the
--  layout of every record is pinned by static_asserts / ctypes size checks on the C side (tests/test_abi.py).
--
--  The first six subprograms are the ones scene_hydra_embree.adb:37-66 already imports from cpp/embree_connect.cpp;
--  their profiles are unchanged, so that body links against libgcore_b200.so without edits.

with System;        use System;
with Interfaces;    use Interfaces;
with Interfaces.C;  use Interfaces.C;
with Interfaces.C.Strings;

package GCore_B200 is

  ------------------------------------------------------------------ legacy seam (cpp/embree_connect.cpp)
  procedure gcore_init_and_clear;
  pragma Import (C, gcore_init_and_clear, "gcore_init_and_clear");

  procedure gcore_destroy;
  pragma Import (C, gcore_destroy, "gcore_destroy");

  function gcore_add_mesh_3f (a_vertices3f : Address; a_vertexNum  : Integer;
                              a_indices    : Address; a_indicesNum : Integer) return Integer;
  pragma Import (C, gcore_add_mesh_3f, "gcore_add_mesh_3f");

  procedure gcore_instance_meshes (a_geomId : Integer; a_matrices : Address; a_matrixNum : Integer);
  pragma Import (C, gcore_instance_meshes, "gcore_instance_meshes");

  procedure gcore_commit_scene;
  pragma Import (C, gcore_commit_scene, "gcore_commit_scene");

  --  36 bytes, the C layout of embree_connect.cpp:186-194 (the record of scene_hydra_embree.adb:53-63 declares one
  --  float too many in both arrays; use this one)
  type Float_3 is array (0 .. 2) of Float;  pragma Convention (C, Float_3);
  type Float_2 is array (0 .. 1) of Float;  pragma Convention (C, Float_2);
  type HitCpp is record
    primIndex : Integer := -1;
    geomIndex : Integer := -1;
    instIndex : Integer := -1;
    t         : Float   := 0.0;
    normal    : Float_3;
    texCoord  : Float_2;
  end record;
  pragma Convention (C, HitCpp);

  function gcore_closest_hit (a_rayPos : Address; a_rayDir : Address; t_near : Float; t_far : Float;
                              pHit : Address) return Interfaces.C.char;
  pragma Import (C, gcore_closest_hit, "gcore_closest_hit");

  ------------------------------------------------------------------ scene of the default (analytic) build
  --  Scene.Init of the internal_ada build on the device: Cornell box, spheres, sphere light, materials, camera,
  --  pyramid2.vsgf under T*R*S (scene.adb:89-217)
  function gcore_scene_init_cornell (vsgf_path : Interfaces.C.Strings.chars_ptr) return int;
  pragma Import (C, gcore_scene_init_cornell, "gcore_scene_init_cornell");

  type Gcore_Ray is record
    origin    : Float_3;
    direction : Float_3;
  end record;
  pragma Convention (C, Gcore_Ray);

  type Gcore_Hit is record                                      --  geometry.ads:57-67
    is_hit, prim_type, prim_index, matId, mat : Integer_32;
    t      : Float;
    normal : Float_3;
    tx, ty : Float;
  end record;
  pragma Convention (C, Gcore_Hit);

  function gcore_closest_hit_batch (n : int; rays : Address; hits : Address) return int;
  pragma Import (C, gcore_closest_hit_batch, "gcore_closest_hit_batch");

  ------------------------------------------------------------------ extension of the seam (scene_hydra_embree.adb)
  --  material id / vertex normals of a mesh added with gcore_add_mesh_3f (call before gcore_commit_scene), and the switch that
  --  makes the integrators trace the committed scene instead of the analytic build's mesh
  function gcore_set_mesh_material (meshId : int; matId : int) return int;
  pragma Import (C, gcore_set_mesh_material, "gcore_set_mesh_material");
  function gcore_set_mesh_normals_3f (meshId : int; a_normals3f : Address; a_vertexNum : int) return int;
  pragma Import (C, gcore_set_mesh_normals_3f, "gcore_set_mesh_normals_3f");
  function gcore_use_committed_scene (enable : int) return int;
  pragma Import (C, gcore_use_committed_scene, "gcore_use_committed_scene");
  --  0: the reference's one-sided triangle test (default), 1: Embree's two-sided rtcIntersect1 semantics
  function gcore_set_seam_two_sided (enable : int) return int;
  pragma Import (C, gcore_set_seam_two_sided, "gcore_set_seam_two_sided");

  ------------------------------------------------------------------ rendering
  type Gcore_Render_Config is record                            --  ray_tracer.ads:20-32,40
    render_type   : Integer_32;                                 --  Render_Type'Pos
    anti_aliasing : Integer_32;
    max_depth     : Integer_32;
    seed          : Unsigned_32;
    background    : Float_3;
  end record;
  pragma Convention (C, Gcore_Render_Config);

  type Gcore_Render_Stats is record
    rays, samples, mesh_queries, bounces, reserved0 : Unsigned_64;
    kernel_ms, total_ms : Float;
    spp, kernel_launches : Integer_32;
  end record;
  pragma Convention (C, Gcore_Render_Stats);

  function gcore_set_device (device : int) return int;
  pragma Import (C, gcore_set_device, "gcore_set_device");
  --  multi-GPU inside this one process: devices [device, device + n) split the tiles of every pass, gcore_read_frame gathers the
  --  accumulators on the first device over NCCL (include/gcore.h); call once before Resize_Viewport
  function gcore_set_gpu_count (n : int) return int;
  pragma Import (C, gcore_set_gpu_count, "gcore_set_gpu_count");
  function gcore_init_render (cfg : access constant Gcore_Render_Config) return int;
  pragma Import (C, gcore_init_render, "gcore_init_render");
  function gcore_resize_viewport (width, height : int) return int;
  pragma Import (C, gcore_resize_viewport, "gcore_resize_viewport");
  function gcore_render_units (units : int; stats : access Gcore_Render_Stats) return int;
  pragma Import (C, gcore_render_units, "gcore_render_units");
  function gcore_get_spp return int;
  pragma Import (C, gcore_get_spp, "gcore_get_spp");
  --  hdr4 : float4 per pixel or Null_Address; ldr : Unsigned_32 per pixel, ROW-major [y*W + x]
  function gcore_read_frame (hdr4 : Address; ldr : Address) return int;
  pragma Import (C, gcore_read_frame, "gcore_read_frame");
  function gcore_last_error (buf : Address; n : int) return int;
  pragma Import (C, gcore_last_error, "gcore_last_error");

end GCore_B200;
