--  ray_tracer-render_pass_b200.adb -- the three subprogram bodies of ray_tracer.adb that change when the render hot
--  path runs on the GPU.  NOT COMPILED HERE (no GNAT in the image).  Everything else of Ray_Tracer (the spec
--  ray_tracer.ads:20-48 included) stays as it is: width/height/Threads_Num/Anti_Aliasing_On/Max_Trace_Depth/
--  Background_Color keep their meaning, screen_buffer keeps its type and its (x, y) indexing.
--
--  with GCore_B200; use GCore_B200;   -- added to the context clause of ray_tracer.adb

  procedure Init_Render(a_rendType : Render_Type) is            -- replaces ray_tracer.adb:197-200
  begin
    g_rend_type := a_rendType;
  end Init_Render;

  procedure Resize_Viewport(size_x, size_y : integer) is        -- replaces ray_tracer.adb:297-320
    rc : Interfaces.C.int;
  begin
    width  := Positive(size_x);
    height := Positive(size_y);
    delete(screen_buffer);
    screen_buffer := new ScreenBufferData(0..width-1, 0..height-1);
    rc := gcore_resize_viewport(Interfaces.C.int(width), Interfaces.C.int(height));   -- accumulator := 0, spp := 0 on the device
    pragma Assert(rc = 0);
  end Resize_Viewport;

  procedure Render_Pass is                                      -- replaces ray_tracer.adb:240-293 and the task pool :142-194
    cfg   : aliased Gcore_Render_Config;
    stats : aliased Gcore_Render_Stats;
    rc    : Interfaces.C.int;
    -- row-major staging buffer: the library writes ldr(y*width + x); screen_buffer is (x, y)-indexed (ray_tracer.ads:35)
    type Row_Major is array (0 .. width*height - 1) of Unsigned_32;
    pragma Convention (C, Row_Major);
    ldr   : Row_Major;
  begin
    cfg.render_type   := Integer_32(Render_Type'Pos(g_rend_type));
    cfg.anti_aliasing := (if Anti_Aliasing_On then 1 else 0);
    cfg.max_depth     := Integer_32(Max_Trace_Depth);
    cfg.seed          := 16#5EED0003#;
    cfg.background    := (Background_Color.x, Background_Color.y, Background_Color.z);
    rc := gcore_init_render(cfg'Access);                                         pragma Assert(rc = 0);
    -- one unit = one Path_Trace_Thread doing one DoPass: Threads_Num units per pass, 4 (AA) or 1 spp each
    rc := gcore_render_units(Interfaces.C.int(Threads_Num), stats'Access);       pragma Assert(rc = 0);
    rc := gcore_read_frame(System.Null_Address, ldr'Address);                    pragma Assert(rc = 0);
    for y in 0 .. height - 1 loop
      for x in 0 .. width - 1 loop
        screen_buffer(x,y) := ldr(y*width + x);
      end loop;
    end loop;
    if g_rend_type = RT_DEBUG or g_rend_type = RT_WHITTED then
      g_finish := true;                                                          -- ray_tracer.adb:259
    end if;
  end Render_Pass;

  function GetSPP return integer is                             -- replaces ray_tracer.adb:322-325
  begin
    return Integer(gcore_get_spp);
  end GetSPP;
