"""GPU (-m gpu): DenseEngine.to_device -- the page-locked upload ring (pageable arrays >= 1 MB) and the direct paths
(small arrays, dtype conversions, arrays that are page-locked already) all deliver the caller's bytes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_upload_ring_delivers_the_bytes_through_slot_reuse_and_growth(engine):
    from b200dense.engine import UploadRing
    rng = np.random.default_rng(0)
    st = torch.cuda.Stream(device=engine.device)
    shapes = [(1200, 1600), (97, 131), (2160, 3840, 3), (1200, 1600), (1080, 1920), (2160, 3840, 3), (600, 3000),
              (1200, 1600), (4000, 3000, 3), (1200, 1600), (1200, 1601), (1 << 20,), (1 << 18,)]
    host, dev = [], []
    with torch.cuda.stream(st):                    # more uploads in a row than the ring has slots; sizes grow and shrink
        for k, shp in enumerate(shapes):
            if len(shp) == 3 or k % 4 == 3:
                a, dt = rng.integers(0, 256, size=shp).astype(np.uint8), torch.uint8
            else:
                a, dt = rng.random(shp, dtype=np.float32), torch.float32
            host.append(a)
            dev.append(engine.to_device(a, dt))
    st.synchronize()
    assert getattr(engine, '_upload_ring', None) is not None and engine._upload_ring._next > UploadRing.SLOTS
    for a, d in zip(host, dev):
        assert d.shape == a.shape and np.array_equal(d.cpu().numpy(), a)


def test_direct_paths_bypass_the_ring(engine):
    rng = np.random.default_rng(1)
    before = engine._upload_ring._next if getattr(engine, '_upload_ring', None) is not None else 0
    small = rng.random((64, 48), dtype=np.float32)                                   # below 1 MB
    f64 = rng.random((600, 800))                                                      # float64 -> float32 conversion
    pinned = torch.from_numpy(rng.random((1200, 1600), dtype=np.float32)).pin_memory()
    strided = rng.random((1200, 3200), dtype=np.float32)[:, ::2]                      # made contiguous first
    d_small, d_f64 = engine.to_device(small), engine.to_device(f64)
    d_pin = engine.to_device(pinned.numpy())
    after_direct = engine._upload_ring._next if getattr(engine, '_upload_ring', None) is not None else 0
    d_str = engine.to_device(strided)
    torch.cuda.synchronize()
    assert after_direct == before                                                     # none of the three used a slot
    assert np.array_equal(d_small.cpu().numpy(), small)
    assert np.array_equal(d_f64.cpu().numpy(), f64.astype(np.float32)) and d_f64.dtype == torch.float32
    assert np.array_equal(d_pin.cpu().numpy(), pinned.numpy())
    assert np.array_equal(d_str.cpu().numpy(), np.ascontiguousarray(strided))


def test_ring_can_be_disabled(engine, monkeypatch):
    monkeypatch.setenv('B200D_UPLOAD', 'direct')
    a = np.random.default_rng(2).integers(0, 256, size=(1080, 1920, 3)).astype(np.uint8)
    before = engine._upload_ring._next if getattr(engine, '_upload_ring', None) is not None else 0
    d = engine.to_device(a, torch.uint8)
    torch.cuda.synchronize()
    assert (engine._upload_ring._next if getattr(engine, '_upload_ring', None) is not None else 0) == before
    assert np.array_equal(d.cpu().numpy(), a)
