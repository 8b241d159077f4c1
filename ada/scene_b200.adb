--  scene_b200.adb -- a third body of package Scene (spec scene.ads unchanged), selected with SCN="b200" in art_b200.gpr.
--  NOT COMPILED HERE (no GNAT in the image).  The scene description lives on the device; the Ada side keeps what the
--  unchanged spec promises (Material_At / Light_At / Camera_At answer from a host copy built like scene.adb does).

with System;
with Interfaces.C;          use Interfaces.C;
with Interfaces.C.Strings;  use Interfaces.C.Strings;
with GCore_B200;            use GCore_B200;

package body Scene is

  procedure Init(a_scn : in out Render_Scene; a_path : in String) is     -- replaces scene.adb:24-27, 89-217
    path : chars_ptr := New_String("../data/pyramid2.vsgf");                -- the relative path scene.adb:206 hard-codes
    rc   : int;
  begin
    rc := gcore_scene_init_cornell(path);
    Free(path);
    pragma Assert(rc = 0);
    Init_Host_Copy(a_scn);      -- the material / light / camera tables of scene.adb:102-192, 212-215, kept for *_At
  end Init;

  procedure Destroy(a_scn : in out Render_Scene) is
  begin
    null;
  end Destroy;

  --  One ray per call is only for callers outside the render loop (picking, tests); the render loop itself runs on the
  --  device behind Ray_Tracer.Render_Pass and never comes through here.
  function Find_Closest_Hit(a_scn : in Render_Scene; r : in Geometry.Ray) return Geometry.Hit is   -- replaces scene.adb:56-86
    ray : aliased Gcore_Ray := (origin    => (r.origin.x, r.origin.y, r.origin.z),
                                direction => (r.direction.x, r.direction.y, r.direction.z));
    h   : aliased Gcore_Hit;
    rc  : int;
    res : Geometry.Hit;
  begin
    rc := gcore_closest_hit_batch(1, ray'Address, h'Address);
    pragma Assert(rc = 0);
    res.is_hit     := (h.is_hit /= 0);
    res.prim_type  := Geometry.Primitive_Type'Val(Integer(h.prim_type));
    res.prim_index := Integer(h.prim_index);
    res.t          := h.t;
    res.normal     := (h.normal(0), h.normal(1), h.normal(2));
    res.matId      := Integer(h.matId);
    res.mat        := a_scn.materials(Integer(h.mat));
    res.tx         := h.tx;
    res.ty         := h.ty;
    return res;
  end Find_Closest_Hit;

  function Material_At(a_scn : in Render_Scene; id : in Integer) return Materials.MaterialRef is (a_scn.materials(id));
  function Light_At   (a_scn : in Render_Scene; id : in Integer) return Lights.LightRef is (a_scn.g_lightRef);
  function Camera_At  (a_scn : in Render_Scene; id : in Integer) return Camera is (a_scn.g_cam);

end Scene;
