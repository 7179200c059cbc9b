"""The MoveIt plugin sources (plugin/moveit_collision_detection_b200) cannot be built here - no MoveIt, Eigen, pluginlib
or ROS in this image - so they are compiled with `g++ -fsyntax-only -Wall -Wextra -Werror` against tests/moveit_stubs/:
declarations-only headers that repeat the members of the reference headers the plugin uses (each with its file:line).
That checks every call the adapter makes into the tested host mirror (moveit_b200/host) and every override against the
CollisionEnv / CollisionDetectorAllocator / CollisionPlugin signatures."""
import os
import shutil
import subprocess
import xml.etree.ElementTree as ET

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "plugin", "moveit_collision_detection_b200")
SOURCES = ["src/collision_env_b200.cpp", "src/collision_detector_b200_plugin_loader.cpp"]


@pytest.mark.parametrize("src", SOURCES)
def test_plugin_source_compiles_against_the_reference_signatures(src):
    gxx = shutil.which("g++")
    assert gxx, "g++ is part of the image"
    cmd = [gxx, "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only",
           "-I" + os.path.join(ROOT, "tests", "moveit_stubs"), "-I" + os.path.join(PLUGIN, "include"),
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "moveit_b200", "host"),
           os.path.join(PLUGIN, src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


def test_plugin_description_names_the_loader_and_the_build_lists_the_sources():
    lib = ET.parse(os.path.join(PLUGIN, "collision_detector_b200_description.xml")).getroot()
    assert lib.tag == "library" and lib.attrib["path"] == "libcollision_detector_b200_plugin"
    cls = lib.find("class")
    assert cls.attrib["name"] == "B200DistanceField"
    assert cls.attrib["type"] == "collision_detection::CollisionDetectorB200PluginLoader"
    assert cls.attrib["base_class_type"] == "collision_detection::CollisionPlugin"
    loader = open(os.path.join(PLUGIN, "src", "collision_detector_b200_plugin_loader.cpp")).read()
    assert "PLUGINLIB_EXPORT_CLASS(%s, %s)" % (cls.attrib["type"], cls.attrib["base_class_type"]) in loader
    cmake = open(os.path.join(PLUGIN, "CMakeLists.txt")).read()
    for src in SOURCES:
        assert src in cmake
    assert "collision_detector_b200_description.xml" in open(os.path.join(PLUGIN, "package_export.xml")).read()
    # the allocator name the description advertises is the one the allocators report
    alloc = open(os.path.join(PLUGIN, "include", "moveit", "collision_detection_b200",
                              "collision_detector_allocator_b200.h")).read()
    assert 'NAME("B200DistanceField")' in alloc
