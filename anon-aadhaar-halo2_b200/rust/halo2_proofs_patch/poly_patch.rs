// Replacement for the `Polynomial` type of halo2_proofs/src/poly.rs @ v2023_01_20.
// SOURCE ONLY (no Rust toolchain in this image).
//
// Public surface unchanged: `Polynomial<F, B>` still derefs to `[F]`, indexes, iterates and clones
// as upstream.  One private field is added: whether the library may hold a device mirror of
// `values` (include/b200zk.h, b200zk_mirror_*).  It is set by the fork's own call sites when they
// hand the buffer to a host-pointer entry point, and every path by which safe code can change or
// release the buffer clears it and tells the library first:
//   * `DerefMut` / `IndexMut` / `iter_mut` (all go through `deref_mut`),
//   * `Drop`,
//   * crate-internal code that takes `values` by value goes through `into_values()`.
// A `&Polynomial` can only be read, so a mirror made from it stays equal to it for as long as the
// flag is set: that is the contract b200zk_mirror_enable states, enforced here by the borrow checker.
use std::cell::Cell;
use std::marker::PhantomData;
use std::ops::{Deref, DerefMut};

use b200zk_sys as ffi;

pub struct Polynomial<F, B> {
    pub(crate) values: Vec<F>,
    pub(crate) _marker: PhantomData<B>,
    mirrored: Cell<bool>,
}

impl<F, B> Polynomial<F, B> {
    pub(crate) fn from_values(values: Vec<F>) -> Self {
        Polynomial { values, _marker: PhantomData, mirrored: Cell::new(false) }
    }

    /// The fork's call sites (commit_lagrange, lagrange_to_coeff, ...) call this right before they
    /// pass `self.values.as_ptr()` to a host-pointer entry point.
    pub(crate) fn mark_mirrored(&self) {
        self.mirrored.set(true);
    }

    #[inline]
    fn forget_mirror(&self) {
        if self.mirrored.replace(false) {
            unsafe { ffi::b200zk_mirror_invalidate(self.values.as_ptr() as *const core::ffi::c_void, 0) };
        }
    }

    /// Change of basis without touching the data (`lagrange_to_coeff` returns the same buffer under
    /// another marker): the mirror, if any, stays valid and the flag travels with the buffer.
    pub(crate) fn rebase<B2>(mut self) -> Polynomial<F, B2> {
        let mirrored = self.mirrored.replace(false);
        let values = std::mem::take(&mut self.values);
        Polynomial { values, _marker: PhantomData, mirrored: Cell::new(mirrored) }
    }

    /// Ownership of the coefficients leaves the type: the library must forget the buffer.
    pub(crate) fn into_values(mut self) -> Vec<F> {
        self.forget_mirror();
        std::mem::take(&mut self.values)
    }
}

impl<F, B> Deref for Polynomial<F, B> {
    type Target = [F];
    fn deref(&self) -> &[F] {
        &self.values
    }
}

impl<F, B> DerefMut for Polynomial<F, B> {
    fn deref_mut(&mut self) -> &mut [F] {
        self.forget_mirror();            // one predictable branch per mutable access while not mirrored
        &mut self.values
    }
}

impl<F, B> Drop for Polynomial<F, B> {
    fn drop(&mut self) {
        self.forget_mirror();            // the allocator may hand this address to another buffer
    }
}

impl<F: Clone, B> Clone for Polynomial<F, B> {
    fn clone(&self) -> Self {
        Polynomial::from_values(self.values.clone())     // a new buffer: not mirrored
    }
}
