--  scene_b200.adb -- a third body of package Scene (spec scene.ads unchanged), selected with SCN="b200" in art_b200.gpr
--  (ada/art_b200.gpr.patch).  NOT COMPILED IN THIS REPOSITORY (the build image has no GNAT); written as a complete
--  compilation unit against the unchanged specs scene.ads, geometry.ads, materials.ads, lights.ads.
--
--  The scene the integrators trace lives on the device (gcore_scene_init_cornell builds it from the same constants);
--  the Ada side keeps what the unchanged spec promises to its callers: Material_At / Light_At / Camera_At answer from
--  a host copy of the tables of scene.adb:102-192, 212-215, and Find_Closest_Hit forwards single rays to the device.

with System;
with Interfaces.C;          use Interfaces.C;
with Interfaces.C.Strings;  use Interfaces.C.Strings;
with Geometry;              use Geometry;
with Materials;             use Materials;
with Lights;                use Lights;
with GCore_B200;            use GCore_B200;

package body Scene is

  --  Host copy of the constant tables (no geometry is traced on the host: the mesh stays on the device).
  procedure Init_Host_Copy (a_scn : in out Render_Scene) is
    light_pos    : constant float3 := (0.0, 4.5, 1.0);
    light_radius : constant float  := 0.5;
    white        : MaterialRef;
    phong        : MaterialRef;
    glass        : MaterialRef;
  begin
    a_scn.g_lightRef := new SphereLight'(center      => light_pos,                       -- scene.adb:116-128
                                         radius      => light_radius,
                                         intensity   => (10.0, 10.0, 10.0),
                                         surfaceArea => 4.0 * M_PI * light_radius * light_radius);

    if a_scn.materials /= null then
      delete (a_scn.materials);
    end if;
    a_scn.materials := new Materials_Array (0 .. 10);                                    -- scene.adb:153-185
    white := new MaterialLambert'(kd => (0.5, 0.5, 0.5));
    phong := new MaterialPhong'(reflection => (0.75, 0.75, 0.75), cosPower => 80.0);
    glass := new MaterialFresnelDielectric'(reflection   => (0.75, 0.75, 0.75),
                                            transparency => (0.85, 0.85, 0.85),
                                            ior          => 1.75);
    a_scn.materials (0)  := glass;
    a_scn.materials (1)  := white;
    a_scn.materials (2)  := new MaterialLambert'(kd => (0.25, 0.5, 0.0));
    a_scn.materials (3)  := new MaterialLambert'(kd => (0.5, 0.0, 0.0));
    a_scn.materials (4)  := new MaterialLight'(lref => a_scn.g_lightRef);
    a_scn.materials (5)  := new MaterialMirror'(reflection => (0.75, 0.75, 0.75));
    a_scn.materials (8)  := phong;
    a_scn.materials (9)  := white;
    a_scn.materials (10) := white;

    if a_scn.spheres /= null then
      delete (a_scn.spheres);
    end if;
    a_scn.spheres := new Spheres_Array (0 .. 2);                                         -- scene.adb:139-144, 182-192
    a_scn.spheres (0) := (pos => (-1.5, 1.0, 1.5), r => 1.0, mat => phong);
    a_scn.spheres (1) := (pos => (1.4, 1.0, 3.0),  r => 1.0, mat => glass);
    a_scn.spheres (2) := (pos => light_pos, r => light_radius, mat => a_scn.materials (4));

    a_scn.g_cam.pos    := (0.0, 2.55, 12.5);                                             -- scene.adb:212-215
    a_scn.g_cam.lookAt := (0.0, 0.0, 0.0);
    a_scn.g_cam.up     := (0.0, 1.0, 0.0);
    a_scn.g_cam.matrix := IdentityMatrix;
  end Init_Host_Copy;

  procedure Init (a_scn : in out Render_Scene; a_path : in String) is                   -- replaces scene.adb:24-27, 89-217
    pragma Unreferenced (a_path);                                                        -- ignored by the reference as well
    path : chars_ptr := New_String ("../data/pyramid2.vsgf");                            -- the relative path scene.adb:206 hard-codes
    rc   : int;
  begin
    rc := gcore_scene_init_cornell (path);
    Free (path);
    if rc /= 0 then
      raise Program_Error with "gcore_scene_init_cornell failed (gcore_last_error has the text)";
    end if;
    Init_Host_Copy (a_scn);
  end Init;

  procedure Destroy (a_scn : in out Render_Scene) is
  begin
    if a_scn.spheres /= null then
      delete (a_scn.spheres);
    end if;
    if a_scn.materials /= null then
      delete (a_scn.materials);
    end if;
  end Destroy;

  --  One ray per call is for callers outside the render loop (picking, tests); the render loop itself runs on the
  --  device behind Ray_Tracer.Render_Pass and never comes through here.
  function Find_Closest_Hit (a_scn : in Render_Scene; r : in Geometry.Ray) return Geometry.Hit is   -- replaces scene.adb:56-86
    ray : aliased Gcore_Ray := (origin    => (r.origin.x, r.origin.y, r.origin.z),
                                direction => (r.direction.x, r.direction.y, r.direction.z));
    h   : aliased Gcore_Hit;
    rc  : int;
  begin
    rc := gcore_closest_hit_batch (1, ray'Address, h'Address);
    if rc /= 0 then
      raise Program_Error with "gcore_closest_hit_batch failed";
    end if;
    declare
      res : Geometry.Hit (Geometry.Primitive'Val (Integer (h.prim_type)));              -- the discriminant of geometry.ads:57
    begin
      res.is_hit     := (h.is_hit /= 0);
      res.prim_index := Integer (h.prim_index);
      res.t          := h.t;
      res.normal     := (h.normal (0), h.normal (1), h.normal (2));
      res.matId      := Integer (h.matId);
      res.tx         := h.tx;
      res.ty         := h.ty;
      if a_scn.materials /= null and then Integer (h.mat) in a_scn.materials'Range then
        res.mat := a_scn.materials (Integer (h.mat));
      end if;
      return res;
    end;
  end Find_Closest_Hit;

  function Material_At (a_scn : in Render_Scene; id : in Integer) return Materials.MaterialRef is
  begin
    if a_scn.materials /= null and then id in a_scn.materials'Range then
      return a_scn.materials (id);
    end if;
    return null;
  end Material_At;

  function Light_At (a_scn : in Render_Scene; id : in Integer) return Lights.LightRef is
    pragma Unreferenced (id);
  begin
    return a_scn.g_lightRef;
  end Light_At;

  function Camera_At (a_scn : in Render_Scene; id : in Integer) return Camera is
    pragma Unreferenced (id);
  begin
    return a_scn.g_cam;
  end Camera_At;

end Scene;
