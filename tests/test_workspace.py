"""".inv" workspace loader of the host layer: XML structure, topology parameters, property
application.  The fixture tests/golden/mini_workspace.inv is written in the layout of the reference's
workspaces/CorrelatedPhotonMappingSingleVolume.inv; where /root/reference is mounted (this container, not the GPU
box) the reference's own file is read too and checked against the values SURVEY.md cites from it."""
import importlib
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, PKG_NAME

FIXTURE = GOLDEN / "mini_workspace.inv"
REFERENCE_WS = Path("/root/reference/workspaces/CorrelatedPhotonMappingSingleVolume.inv")
TRACER = "org.inviwo.ProgressivePhotonTracerCL"
IMPORTANCE = "org.inviwo.MinMaxUniformGrid3DImportanceCLProcessor"
LIGHTVOL = "org.inviwo.PhotonToLightVolumeProcessorCL"
FIXTURE_TF = [(0.05, (1.0, 0.6, 0.25, 0.0)), (0.3, (0.6, 0.25, 0.125, 0.2)), (0.7, (0.125, 0.625, 0.65, 0.5))]


@pytest.fixture(scope="module")
def host():
    return importlib.import_module(PKG_NAME + ".host")


def test_describe_fixture(host):
    d = host.workspace_describe(FIXTURE)
    types = [p[0] for p in d["processors"]]
    assert types.count("org.inviwo.DirectionalLightSamplerCL") == 2 and types.count("org.inviwo.Directionallightsource") == 2
    by_name = {p[1]: p for p in d["processors"]}
    tracer = by_name["ProgressivePhotonTracer"]
    assert tracer[0] == TRACER
    # only properties that store a value are listed; nested ones by path
    assert set(tracer[2]) == {"radius", "maxScatteringEvents", "material.phaseFunction", "material.anisotropy",
                              "transferFunction", "maxIncrementalPhotonsToUpdate", "clipX", "clipY", "clipZ"}
    assert by_name["Key light"][2] == ["lightPosition.position", "lighting.lightDiffuse", "lighting.lightPower"]
    assert ("MinMaxUniformGrid3DImportance.importanceUniformGrid3D", "ProgressivePhotonTracer.recomputationImportance") in d["connections"]
    assert ("Directional light sampler 2.LightSamples", "ProgressivePhotonTracer.LightSamples") in d["connections"]
    assert len(d["connections"]) == 16 and not any("?" in a or "?" in b for a, b in d["connections"])


def test_config_from_fixture(host):
    cfg = host.workspace_config(FIXTURE, basis=(10.0, 10.0, 20.0))
    assert (cfg.samples_per_side, cfg.n_lights, cfg.max_scattering_events) == (64, 2, 3)
    assert (cfg.light_volume_option, cfg.light_volume_channels, cfg.with_importance_grid) == (2, 4, 1)
    assert cfg.photon_radius_voxels == 1.5 and cfg.max_incremental_percent == 60.0 and cfg.incremental_threshold_percent == 35.0
    assert list(cfg.clip) == [6, 48, 0, 38, 0, 36]
    # key light: positionWorldSpace (-30, 50, -40) -> direction -p / basis, normalised
    d = -np.array([-30.0, 50.0, -40.0]) / np.array([10.0, 10.0, 20.0])
    assert np.allclose(list(cfg.light_directions[0]), d / np.linalg.norm(d), atol=1e-6)
    assert np.allclose(list(cfg.light_intensity[0]), [2.0, 1.0, 0.5])
    # fill light: no world-space position stored -> the position property; default power and colour
    d = -np.array([4.0, 2.0, -1.0]) / np.array([10.0, 10.0, 20.0])
    assert np.allclose(list(cfg.light_directions[1]), d / np.linalg.norm(d), atol=1e-6)
    assert np.allclose(list(cfg.light_intensity[1]), [1.0, 1.0, 1.0])


@pytest.mark.parametrize("text,what", [
    ("", "no root"), ("<a><b></a>", "closes"), ("<a x=1/>", "quoted"), ("<a>", "missing </a>"),
    ("<NotAWorkspace/>", "not an Inviwo workspace"), ("<InviwoTreeData/>", "no <Processors>"),
    ("<InviwoTreeData><Processors/></InviwoTreeData><x/>", "after the root"),
])
def test_malformed_workspaces_are_refused(host, tmp_path, text, what):
    p = tmp_path / "bad.inv"
    p.write_text(text)
    with pytest.raises(host.HostError) as e:
        host.workspace_describe(p)
    assert what in str(e.value)
    with pytest.raises(host.HostError):
        host.workspace_describe(tmp_path / "missing.inv")


def test_out_of_range_property_is_refused_by_config_free_parse(host, tmp_path):
    """values are validated when they are applied to a processor (GPU test below); parsing alone accepts them"""
    text = FIXTURE.read_text().replace('<value content="3" />', '<value content="99" />')
    p = tmp_path / "range.inv"
    p.write_text(text)
    assert host.workspace_config(p).max_scattering_events == 99


@pytest.mark.skipif(not REFERENCE_WS.exists(), reason="reference tree not mounted (GPU box)")
def test_reference_workspace_values(host):
    """the numbers SURVEY.md 8(b)/(d) quote from the reference's workspace"""
    d = host.workspace_describe(REFERENCE_WS)
    types = [p[0] for p in d["processors"]]
    for cid in (TRACER, IMPORTANCE, LIGHTVOL, "org.inviwo.UniformSampleGenerator2DCL", "org.inviwo.VolumeMinMaxCLProcessor"):
        assert types.count(cid) == 1, cid
    assert types.count("org.inviwo.DirectionalLightSamplerCL") == 2
    cfg = host.workspace_config(REFERENCE_WS, basis=(399.2, 399.2, 76.0))
    assert (cfg.samples_per_side, cfg.n_lights, cfg.max_scattering_events) == (1024, 2, 1)      # ws:444-446
    assert cfg.light_volume_option == 2 and cfg.with_importance_grid == 1                        # ws:555-557
    assert list(cfg.clip) == [73, 512, 7, 512, 0, 96]                                            # ws:740-757
    for l in range(2):
        assert abs(np.linalg.norm(list(cfg.light_directions[l])) - 1.0) < 1e-5
    imp = {p[1]: p for p in d["processors"]}["MinMaxUniformGrid3DImportance"]
    assert {"constantWeight", "opacityDiffWeight", "colorWeight", "colorDiffWeight", "transferfunction"} <= set(imp[2])


@pytest.mark.gpu
def test_network_from_workspace_equals_network_configured_by_hand(host, cpm, synth, torch_cuda):
    """every stored property reaches its processor: the network loaded from the file and the network set up through
    the API with the same values trace the same photons and build the same light volume"""
    dims = (48, 40, 36)
    vol = synth.volume_u8(dims, 5)
    a = host.Network.from_workspace(FIXTURE, dims, cpm.CPM_FMT_U8)
    assert a.properties_applied >= 17
    assert a.get_property(TRACER, "maxScatteringEvents") == 3 and a.get_property(TRACER, "radius") == 1.5
    assert a.get_property(TRACER, "maxIncrementalPhotonsToUpdate") == 60.0
    assert a.get_property(TRACER, "transferFunction") == 3                       # number of TF points
    assert a.get_property(IMPORTANCE, "constantWeight") == 0.25 and a.get_property(IMPORTANCE, "opacityDiffWeight") == 0.75
    assert a.get_property(IMPORTANCE, "colorWeight") == 0.0 and a.get_property(IMPORTANCE, "useAssociatedColor") == 1.0
    assert a.get_property(LIGHTVOL, "volumeSizeOption") == 2 and a.get_property(LIGHTVOL, "volumeDataType") == 4
    assert a.get_property(LIGHTVOL, "incrementalRecomputationThreshold") == 35.0
    assert a.get_property("org.inviwo.DirectionalLightSamplerCL", "wgsize", 0) == 128
    a.set_volume_host(vol)
    a.evaluate()
    pa, la = a.read_photons(3), a.read_light_volume()
    cfg = host.workspace_config(FIXTURE)
    b = host.Network(dims, cpm.CPM_FMT_U8, 64, [tuple(cfg.light_directions[i]) for i in range(2)], max_scattering_events=3,
                     light_volume_option=2, light_volume_channels=4, with_importance_grid=True, photon_radius_voxels=1.5,
                     max_incremental_percent=60.0, clip=[6, 48, 0, 38, 0, 36],
                     light_intensity=[tuple(cfg.light_intensity[i]) for i in range(2)], incremental_threshold=35.0)
    b.set_transfer_function(FIXTURE_TF)
    # the material (Henyey-Greenstein, g = 0.4) has no setter in the headless API: taken from the file here too
    assert b.load_workspace(FIXTURE) >= 17
    b.set_volume_host(vol)
    b.evaluate()
    pb, lb = b.read_photons(3), b.read_light_volume()
    assert pa.shape == (2 * 64 * 64 * 3, 8) and (pa[:, 0] < 1e38).sum() > 500
    assert np.array_equal(pa.view(np.uint32), pb.view(np.uint32))
    assert la.shape == lb.shape and np.allclose(la, lb, rtol=1e-5, atol=1e-9) and la.max() > 0
    assert a.get_property(TRACER, "enableProgressiveRecomputation") == 1.0      # not stored: the processor's default
    # an out-of-range stored value is refused when it is applied
    bad = Path(str(FIXTURE)).read_text().replace('<value content="3" />', '<value content="99" />')
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".inv", delete=False) as f:
        f.write(bad)
    with pytest.raises(host.HostError) as e:
        a.load_workspace(f.name)
    assert "maxScatteringEvents" in str(e.value)
    a.close(); b.close()


def test_generated_workspaces_round_trip_through_describe(host, tmp_path):
    """structure check on machine-written files: comments, declarations, single-quoted and escaped attribute values,
    deep property nesting, ports that carry ids on either end of a connection"""
    import random
    rnd = random.Random(7)
    for case in range(20):
        n = rnd.randint(1, 6)
        procs, names = [], []
        for i in range(n):
            name = f"P {i} <&> 'q'"
            esc = name.replace("&", "&amp;").replace("<", "&lt;").replace(">", "&gt;").replace("'", "&apos;")
            names.append(name)
            depth = rnd.randint(0, 4)
            inner = '<Property type="t" identifier="leaf"><value content="1.5" /></Property>'
            for d in range(depth):
                inner = f'<Property type="c" identifier="c{d}"><Properties>{inner}</Properties><!-- c --></Property>'
            procs.append(f"""<Processor type='org.test.T{i % 2}' identifier="{esc}">
                <InPorts><InPort type="x" identifier="in" id="refi{i}" /></InPorts>
                <OutPorts><OutPort type="x" identifier="out" id="refo{i}" /></OutPorts>
                <Properties>{inner}<Property type="b" identifier="flag" /></Properties></Processor>""")
        conns = "".join(f'<Connection><OutPort type="x" identifier="out" reference="refo{i}" />'
                        f'<InPort type="x" identifier="in" reference="refi{i + 1}" /></Connection>' for i in range(n - 1))
        text = ('<?xml version="1.0" ?>\n<!-- generated -->\n<InviwoTreeData version="1.0"><Processors>' + "".join(procs) +
                f"</Processors><Connections>{conns}</Connections></InviwoTreeData>\n")
        p = tmp_path / f"gen{case}.inv"
        p.write_text(text)
        d = host.workspace_describe(p)
        assert [q[1] for q in d["processors"]] == names
        assert len(d["connections"]) == n - 1 and all("?" not in a + b for a, b in d["connections"])
        for q in d["processors"]:
            assert len(q[2]) == 1 and q[2][0].endswith("leaf") and q[2][0].count(".") <= 4
