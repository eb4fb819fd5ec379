// TEST INFRASTRUCTURE -- the fiber switch of tests/emu/cuda_runtime.h for x86-64 (System V): push the callee-saved
// registers, exchange stack pointers, pop, return.  No signal mask, no floating-point environment: every fiber runs on
// the same OS thread with the same settings.
#if defined(__x86_64__)
asm(R"(
    .text
    .globl emu_switch
    .type emu_switch, @function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq (%rsi), %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size emu_switch, .-emu_switch
    .section .note.GNU-stack,"",@progbits
)");
#endif
