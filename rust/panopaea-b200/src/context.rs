//! Device + stream + scratch (`pano_ctx`).  The reference has no such object (it computes on the calling thread and
//! the global rayon pool); `Grid2d::new` therefore uses a lazily created per-thread default context so that the
//! example's `Grid2d::new((128, 128))` keeps working unchanged.
use std::cell::RefCell;
use std::ffi::CString;
use std::ptr;
use std::rc::Rc;

use crate::ffi;

pub(crate) struct Inner {
    pub(crate) raw: *mut ffi::pano_ctx,
}

impl Drop for Inner {
    fn drop(&mut self) {
        unsafe {
            ffi::pano_ctx_destroy(self.raw);
        }
    }
}

/// Cheap to clone; the device context lives as long as any clone, grid or field that refers to it.
#[derive(Clone)]
pub struct Context {
    pub(crate) inner: Rc<Inner>,
}

thread_local! {
    static DEFAULT: RefCell<Option<Context>> = RefCell::new(None);
}

impl Context {
    /// A private stream on `device`.  Panics when no sm_100 device is usable: there is no CPU fallback.
    pub fn new(device: i32) -> Context {
        let mut raw = ptr::null_mut();
        ffi::check(unsafe { ffi::pano_ctx_create(device, ptr::null_mut(), &mut raw) });
        Context { inner: Rc::new(Inner { raw }) }
    }

    /// The calling thread's default context: device `PANOPAEA_B200_DEVICE` (default 0), created on first use.
    pub fn default_for_thread() -> Context {
        DEFAULT.with(|slot| {
            let mut slot = slot.borrow_mut();
            if slot.is_none() {
                let device = std::env::var("PANOPAEA_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
                *slot = Some(Context::new(device));
            }
            slot.as_ref().unwrap().clone()
        })
    }

    pub fn raw(&self) -> *mut ffi::pano_ctx {
        self.inner.raw
    }

    /// Waits for everything enqueued on the context's stream.
    pub fn sync(&self) {
        ffi::check(unsafe { ffi::pano_ctx_sync(self.raw()) });
    }

    /// Tuning knobs of the library (`cg_kernel`, `cg_dynamic`, `step_timing`, ...; see the header).
    pub fn set_option(&self, key: &str, value: i64) {
        let key = CString::new(key).expect("option key");
        ffi::check(unsafe { ffi::pano_ctx_set_option(self.raw(), key.as_ptr(), value) });
    }

    pub fn launch_count(&self) -> u64 {
        let mut n = 0u64;
        ffi::check(unsafe { ffi::pano_ctx_launch_count(self.raw(), &mut n) });
        n
    }

    /// CUDA-event timing on the stream the kernels run on.
    pub fn timer_start(&self) {
        ffi::check(unsafe { ffi::pano_timer_start(self.raw()) });
    }

    pub fn timer_stop_ms(&self) -> f64 {
        let mut ms = 0.0;
        ffi::check(unsafe { ffi::pano_timer_stop_ms(self.raw(), &mut ms) });
        ms
    }
}
