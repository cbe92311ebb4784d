"""GPU (-m gpu): device `_prepare_images` == OpenCV, and reconstruct() from device-prepared images == host-prepared."""
import cv2 as cv
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('h,w,scale', [(97, 131, 0.25), (480, 640, 0.37), (600, 800, 0.5), (72, 96, 1.0), (50, 70, 1.7)])
def test_prepare_image_kernel_matches_opencv(engine, h, w, scale):
    rng = np.random.default_rng(h + w)
    img = rng.integers(0, 256, size=(h, w, 3)).astype(np.uint8)
    nh, nw = int(h * scale), int(w * scale)
    color, gray = engine.prepare_image(engine.to_device(img, torch.uint8), nh, nw)
    ref_color = cv.resize(img, (nw, nh))
    ref_gray = cv.cvtColor(ref_color, cv.COLOR_BGR2GRAY).astype(np.float32) / 255.0
    assert np.array_equal(color.cpu().numpy(), ref_color)
    assert np.array_equal(gray.cpu().numpy(), ref_gray)


def test_reconstruct_device_prepare_equals_host_prepare():
    """reconstruct() fed by the GPU resize/gray kernel returns the same cloud as with host OpenCV preparation."""
    from b200dense.scenes import make_scene
    from core.camera import Camera, CameraPose
    from core.dense_stereo import DenseStereoReconstructor
    from core.mvs_patchmatch import PatchMatchMVS
    sc = make_scene('wavy', 5, 192, 144, seed=17)
    cam = Camera(K=sc.K.copy(), dist=np.zeros(5))
    poses = {i: CameraPose(R=sc.poses[i][0], t=sc.poses[i][1]) for i in sc.poses}
    sparse = np.stack([np.zeros(20), np.zeros(20), np.linspace(4.0, 6.0, 20)], axis=1)
    clouds = []
    for dev_prep in (True, False):
        pm = PatchMatchMVS(cam, scale=0.5, patch_size=5, num_iterations=2, num_samples=2, min_views=2, seed=3)
        pm.device_prepare = dev_prep
        clouds.append(pm.reconstruct(sc.images, poses, sparse_points=sparse))
    assert len(clouds[0][0]) > 50
    assert np.array_equal(clouds[0][0], clouds[1][0]) and np.array_equal(clouds[0][1], clouds[1][1])
    clouds = []
    for dev_prep in (True, False):
        st = DenseStereoReconstructor(cam, scale=0.5, num_depths=16, patch_size=5)
        st.device_prepare = dev_prep
        clouds.append(st.reconstruct(sc.images, poses))
    assert len(clouds[0][0]) > 50
    assert np.array_equal(clouds[0][0], clouds[1][0]) and np.array_equal(clouds[0][1], clouds[1][1])
